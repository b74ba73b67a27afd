"""Import the UNMODIFIED reference python files (read-only, /root/reference) in the build container.

Test infrastructure only.  The reference cannot be imported as-is: it pulls lightning, hydra, diffusers,
conformer, matplotlib, phonemizer ... none of which are installed.  This module installs
inert stand-ins for those packages in `sys.modules` -- plus a restatement of the four diffusers==0.25.0
symbols the hot path touches -- and then imports `matcha.models.matcha_tts`, `matcha.hifigan.*` from where
they lie.  Nothing is copied into this repository's history.  It is used by scripts/make_golden.py (which writes
tests/golden/), tests/test_oracle_vs_reference.py and the CPU arms of bench.py (`--impl reference`, `cpu_baseline`), which
on the GPU box find the staged copy of the same files under baseline/_ref/ (scripts/install_reference.py).
"""
from __future__ import annotations

import contextlib
import os
import sys
import types

import torch
import torch.nn as nn
import torch.nn.functional as F

def _find_root():
    """EMOJIVOICE_REFERENCE, else the read-only tree of the build container, else the staged copy of the path's files that
    scripts/install_reference.py puts under baseline/_ref/ (git-ignored; it travels to the GPU box with gpurun)."""
    staged = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "baseline", "_ref")
    for cand in (os.environ.get("EMOJIVOICE_REFERENCE"), "/root/reference", staged):
        if cand and os.path.isdir(os.path.join(cand, "Matcha-TTS", "matcha", "models")):
            return cand
    return "/root/reference"


REFERENCE_ROOT = _find_root()
_MATCHA_ROOT = os.path.join(REFERENCE_ROOT, "Matcha-TTS")


def available() -> bool:
    return os.path.isdir(os.path.join(_MATCHA_ROOT, "matcha", "models"))


def _mod(name, **attrs):
    m = types.ModuleType(name)
    m.__dict__.update(attrs)
    m.__path__ = []  # behave as a package so that submodule imports resolve through sys.modules
    sys.modules[name] = m
    return m


class _LightningModule(nn.Module):
    def save_hyperparameters(self, *a, **k):
        pass

    def log(self, *a, **k):
        pass


class _RefAttention(nn.Module):
    """diffusers==0.25.0 `Attention` with `AttnProcessor2_0`, self-attention only (call site transformer.py:196-204)."""

    def __init__(self, query_dim, heads=8, dim_head=64, dropout=0.0, bias=False, cross_attention_dim=None,
                 upcast_attention=False, **_):
        super().__init__()
        inner = heads * dim_head
        self.heads = heads
        self.to_q = nn.Linear(query_dim, inner, bias=bias)
        self.to_k = nn.Linear(cross_attention_dim or query_dim, inner, bias=bias)
        self.to_v = nn.Linear(cross_attention_dim or query_dim, inner, bias=bias)
        self.to_out = nn.ModuleList([nn.Linear(inner, query_dim), nn.Dropout(dropout)])

    def forward(self, hidden_states, encoder_hidden_states=None, attention_mask=None, **_):
        b, t, _c = hidden_states.shape
        ctx = hidden_states if encoder_hidden_states is None else encoder_hidden_states
        q, k, v = self.to_q(hidden_states), self.to_k(ctx), self.to_v(ctx)
        hd = q.shape[-1] // self.heads
        if attention_mask is not None:
            if attention_mask.shape[0] < b * self.heads:
                attention_mask = attention_mask.repeat_interleave(self.heads, dim=0)
            attention_mask = attention_mask.view(b, self.heads, -1, attention_mask.shape[-1])
        q = q.view(b, -1, self.heads, hd).transpose(1, 2)
        k = k.view(b, -1, self.heads, hd).transpose(1, 2)
        v = v.view(b, -1, self.heads, hd).transpose(1, 2)
        o = F.scaled_dot_product_attention(q, k, v, attn_mask=attention_mask, dropout_p=0.0, is_causal=False)
        o = o.transpose(1, 2).reshape(b, -1, self.heads * hd).to(q.dtype)
        return self.to_out[1](self.to_out[0](o))


class _LoRACompatibleLinear(nn.Linear):
    def forward(self, x, scale: float = 1.0):  # no LoRA layer attached == nn.Linear
        return super().forward(x)


_installed = False


def install():
    global _installed
    if _installed:
        return
    if not available():
        raise RuntimeError(f"reference tree not found under {REFERENCE_ROOT}")
    ident = lambda f=None, *a, **k: f if callable(f) else (lambda g: g)
    _mod("lightning", LightningModule=_LightningModule, LightningDataModule=object, Callback=object, Trainer=object)
    _mod("lightning.pytorch", LightningModule=_LightningModule, LightningDataModule=object, Callback=object,
         Trainer=object)
    _mod("lightning.pytorch.utilities", grad_norm=lambda *a, **k: {}, rank_zero_only=ident)
    _mod("lightning.pytorch.loggers", Logger=object)
    _mod("hydra")
    _mod("hydra.core")
    _mod("hydra.core.hydra_config", HydraConfig=object)
    _mod("omegaconf", DictConfig=dict, OmegaConf=object, open_dict=contextlib.nullcontext)
    for name in ("gdown", "wget", "rootutils", "phonemizer", "unidecode", "misaki", "piper_phonemize"):
        _mod(name)
    _mod("matplotlib", use=lambda *a, **k: None)
    _mod("matplotlib.pyplot")
    _mod("matplotlib.pylab")
    _mod("conformer", ConformerBlock=nn.Module)
    _mod("diffusers")
    _mod("diffusers.models")
    _mod("diffusers.utils")
    _mod("diffusers.models.attention", GEGLU=nn.Module, GELU=nn.Module, AdaLayerNorm=nn.Module,
         AdaLayerNormZero=nn.Module, ApproximateGELU=nn.Module)
    _mod("diffusers.models.attention_processor", Attention=_RefAttention)
    _mod("diffusers.models.lora", LoRACompatibleLinear=_LoRACompatibleLinear)
    _mod("diffusers.utils.torch_utils", maybe_allow_in_graph=lambda cls: cls)
    _mod("diffusers.models.activations", get_activation=lambda name: {"silu": nn.SiLU(), "swish": nn.SiLU(),
                                                                        "mish": nn.Mish(), "gelu": nn.GELU()}[name])
    sys.path.insert(0, _MATCHA_ROOT)
    # matcha/utils/__init__.py drags the whole training scaffold in; give the package a minimal face instead.
    import importlib.util
    import logging

    pkg = _mod("matcha")
    pkg.__path__ = [os.path.join(_MATCHA_ROOT, "matcha")]
    utils = _mod("matcha.utils", get_pylogger=lambda name=__name__: logging.getLogger(name))
    utils.__path__ = [os.path.join(_MATCHA_ROOT, "matcha", "utils")]
    _mod("matcha.utils.pylogger", get_pylogger=lambda name=__name__: logging.getLogger(name))
    _install_monotonic_align()
    spec = importlib.util.spec_from_file_location(
        "matcha.utils.model", os.path.join(_MATCHA_ROOT, "matcha", "utils", "model.py"))
    model_mod = importlib.util.module_from_spec(spec)
    sys.modules["matcha.utils.model"] = model_mod
    spec.loader.exec_module(model_mod)
    utils.model = model_mod
    _installed = True


def _install_monotonic_align():
    """MatchaTTS.forward needs the alignment search.  With the reference's own Cython kernel compiled under oracle/_ref
    (oracle/build_oracle.py) the reference's UNMODIFIED wrapper monotonic_align/__init__.py runs on top of it; otherwise the
    C restatement stands in (oracle/mas_oracle.py, itself pinned against the compiled reference by tests/test_mas.py)."""
    import importlib.util

    from . import build_oracle, mas_oracle

    core = build_oracle.load_ref() if build_oracle.ref_module_path() else None
    init_py = os.path.join(_MATCHA_ROOT, "matcha", "utils", "monotonic_align", "__init__.py")
    if core is not None and os.path.exists(init_py):
        sys.modules["matcha.utils.monotonic_align.core"] = core
        spec = importlib.util.spec_from_file_location("matcha.utils.monotonic_align", init_py)
        mod = importlib.util.module_from_spec(spec)
        sys.modules["matcha.utils.monotonic_align"] = mod
        spec.loader.exec_module(mod)
    else:
        _mod("matcha.utils.monotonic_align", maximum_path=mas_oracle.maximum_path)


def build_matcha(cfg, state_dict):
    """Instantiate the reference MatchaTTS with `cfg` and load `state_dict` strictly."""
    install()
    from matcha.models.matcha_tts import MatchaTTS  # noqa: the reference's own class

    model = MatchaTTS(**cfg.constructor_kwargs())
    model.load_state_dict(state_dict, strict=True)
    return model.eval()


def build_hifigan(h, state_dict, checkpoint_form=False):
    """Reference HiFi-GAN generator; `checkpoint_form` loads weight_g/weight_v keys before remove_weight_norm."""
    install()
    import io

    from matcha.hifigan.env import AttrDict
    from matcha.hifigan.models import Generator

    gen = Generator(AttrDict(dict(h)))
    with contextlib.redirect_stdout(io.StringIO()):
        if checkpoint_form:
            gen.load_state_dict(state_dict, strict=True)
            gen.eval()
            gen.remove_weight_norm()
        else:
            gen.eval()
            gen.remove_weight_norm()
            gen.load_state_dict(state_dict, strict=True)
    return gen


def build_denoiser(vocoder):
    install()
    from matcha.hifigan.denoiser import Denoiser

    return Denoiser(vocoder, mode="zeros")


@contextlib.contextmanager
def injected_noise(z: torch.Tensor):
    """Make the reference's `torch.randn_like(mu)` (flow_matching.py:51) return `z` for the duration."""
    orig = torch.randn_like

    def fake(t, *a, **k):
        assert t.shape == z.shape, (t.shape, z.shape)
        return z.to(t.dtype).clone()

    torch.randn_like = fake
    try:
        yield
    finally:
        torch.randn_like = orig
