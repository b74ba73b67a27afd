"""GPU parity tests (run on the B200 box with -m gpu): the CUDA path, called through the C ABI, against the CPU
oracle on the same seeded inputs and against the fixtures the reference itself produced (tests/golden/).

Tolerances (BASELINE.json north_star): durations / alignment bit-exact; mel / waveform rel-L2 <= 1e-4 in fp32 mode and
<= 1e-2 in bf16 mode (tcgen05 operands, fp32 accumulation)."""
import math

import numpy as np
import pytest

import torch
import torch.nn.functional as F

import emojivoice_b200 as ev
from emojivoice_b200 import _lib, synthetic
from emojivoice_b200.config import HIFIGAN_V1, VCTK
from oracle import hifigan_oracle as ho
from oracle import matcha_oracle as mo
from oracle.metrics import mel_cepstral_distortion
from tests import golden_io
from tests.conftest import rel_l2

pytestmark = pytest.mark.gpu

TOL = {"fp32": 1e-4, "bf16": 1e-2}


@pytest.fixture(scope="module")
def ctx():
    return _lib.Context()


@pytest.fixture(scope="module")
def matcha(matcha_sd):
    m = ev.MatchaTTS(**VCTK.constructor_kwargs())
    m.load_state_dict(matcha_sd)
    return m


@pytest.fixture(scope="module")
def vocoders():
    out = {}
    for name, kw in golden_io.HIFIGAN.items():
        sd = synthetic.hifigan_state_dict(HIFIGAN_V1, **kw)
        g = ev.Generator(HIFIGAN_V1)
        g.load_state_dict(sd)
        g.remove_weight_norm()
        out[name] = (g, sd)
    return out


# ------------------------------------------------------------------------------------------------ single kernels
CONV_CASES = [
    # B, Cin, T, Cout, K, stride, pad, dil, transposed
    (2, 64, 200, 64, 1, 1, 0, 1, False), (2, 256, 300, 256, 3, 1, 1, 1, False), (1, 224, 130, 256, 3, 1, 1, 1, False),
    (2, 80, 77, 512, 7, 1, 3, 1, False), (2, 32, 1000, 32, 11, 1, 25, 5, False), (2, 128, 500, 128, 7, 1, 9, 3, False),
    (2, 256, 128, 256, 3, 2, 1, 1, False), (2, 256, 64, 256, 4, 2, 1, 1, True), (2, 512, 50, 256, 16, 8, 4, 1, True),
    (1, 64, 300, 32, 4, 2, 1, 1, True), (2, 256, 100, 80, 1, 1, 0, 1, False), (1, 1024, 200, 256, 1, 1, 0, 1, False),
    (2, 256, 140, 384, 1, 1, 0, 1, False), (1, 256, 1, 256, 3, 1, 1, 1, False), (3, 64, 129, 64, 3, 1, 1, 1, False),
    # two m-blocks per tile + two TMA boxes per haloed tile (BN = 128 / 64 / 32), 64-byte swizzle (C_in <= 32), ragged tails
    (2, 128, 700, 128, 11, 1, 25, 5, False), (1, 64, 900, 64, 7, 1, 9, 3, False), (2, 32, 1500, 32, 3, 1, 1, 1, False),
    (1, 16, 300, 32, 5, 1, 2, 1, False), (1, 32, 257, 64, 11, 1, 5, 1, False), (4, 256, 334, 256, 3, 1, 1, 1, False),
]


@pytest.mark.parametrize("prec", ["fp32", "bf16"])
@pytest.mark.parametrize("case", CONV_CASES)
def test_conv1d_kernel_matches_torch(ctx, case, prec):
    B, Cin, T, Cout, K, stride, pad, dil, transposed = case
    g = torch.Generator().manual_seed(sum(int(v) * (i + 1) for i, v in enumerate(case)))
    x = torch.randn(B, Cin, T, generator=g)
    w = torch.randn((Cin, Cout, K) if transposed else (Cout, Cin, K), generator=g) / (Cin * K) ** 0.5
    b = torch.randn(Cout, generator=g)
    # the tensor-core path rounds its operands to bf16 and accumulates in fp32: compare like for like, tightly
    xr, wr = (x.bfloat16().float(), w.bfloat16().float()) if prec == "bf16" else (x, w)
    if transposed:
        ref = F.conv_transpose1d(xr.double(), wr.double(), b.double(), stride=stride, padding=pad)
    else:
        ref = F.conv1d(xr.double(), wr.double(), b.double(), stride=stride, padding=pad, dilation=dil)
    y = torch.empty(ref.shape, device="cuda")
    xd, wd, bd = x.cuda(), w.cuda(), b.cuda()          # keep the device tensors alive across the call
    rc = _lib.lib().ev_test_conv1d(ctx.handle, _lib.ptr(xd), _lib.ptr(wd), _lib.ptr(bd), B, Cin, T, Cout, K,
                                   stride, pad, dil, int(transposed), _lib.PREC[prec], _lib.ptr(y), _lib.stream_ptr())
    ctx.check(rc, "ev_test_conv1d")
    assert rel_l2(y.cpu(), ref) < 2e-6


TF32_CASES = [  # B, Cin, T, Cout, K, pad, dil -- the text encoder's layer shapes (text_encoder.py) and ragged variants
    (2, 192, 181, 192, 5, 2, 1), (2, 256, 181, 768, 1, 0, 1), (3, 256, 77, 768, 3, 1, 1), (2, 768, 181, 256, 3, 1, 1),
    (1, 256, 300, 80, 1, 0, 1), (2, 256, 130, 256, 3, 1, 1), (1, 64, 9, 160, 3, 1, 1),
]


@pytest.mark.parametrize("case", TF32_CASES)
def test_conv1d_3xtf32_tensor_core_path_is_fp32_accurate(ctx, case):
    """The text encoder's convs run on tcgen05 as x_hi*w_hi + x_hi*w_lo + x_lo*w_hi (tf32 operands) with two-level
    accumulation: the tensor core's truncating fp32 accumulator is flushed into a rounded fp32 master sum every <= 16
    MMAs.  The result must be as close to the exact product as the fp32 CUDA-core kernel's (durations feed a ceil())."""
    B, Cin, T, Cout, K, pad, dil = case
    g = torch.Generator().manual_seed(sum(case))
    x = torch.randn(B, Cin, T, generator=g)
    w = torch.randn(Cout, Cin, K, generator=g) / (Cin * K) ** 0.5
    b = torch.randn(Cout, generator=g)
    ref = F.conv1d(x.double(), w.double(), b.double(), padding=pad, dilation=dil)
    xd, wd, bd = x.cuda(), w.cuda(), b.cuda()
    err = {}
    for prec in ("tf32x3", "fp32"):
        y = torch.empty(ref.shape, device="cuda")
        ctx.check(_lib.lib().ev_test_conv1d(ctx.handle, _lib.ptr(xd), _lib.ptr(wd), _lib.ptr(bd), B, Cin, T, Cout, K, 1, pad, dil, 0,
                                            _lib.PREC[prec], _lib.ptr(y), _lib.stream_ptr()), "ev_test_conv1d")
        err[prec] = rel_l2(y.cpu(), ref)
    assert err["fp32"] < 1e-6
    assert err["tf32x3"] < 1e-6, err


@pytest.mark.parametrize("prec", ["fp32", "bf16"])
@pytest.mark.parametrize("case", [(2, 200, 2, 0), (3, 77, 2, 1), (1, 668, 2, 0), (2, 64, 1, 0), (2, 129, 2, 1), (1, 1, 2, 0)])
def test_decoder_attention_matches_torch_sdpa(ctx, case, prec):
    """diffusers Attention semantics: the 0/1 mask is ADDED to the logits, padded keys stay in the softmax."""
    B, T, H, shift = case
    g = torch.Generator().manual_seed(B * 1000 + T)
    qkv = torch.randn(B, 3 * H * 64, T, generator=g) * 1.5
    lens = torch.randint(1, (T << shift) + 1, (B,), generator=g)
    lens[0] = T << shift
    if B > 2:
        lens[1], lens[2] = 1, max(1, (T << shift) // 3)     # whole tiles of padding: zero-filled, not computed
    if prec == "bf16":
        qkv = qkv.bfloat16().float()
    q, k, v = (t.reshape(B, H, 64, T).transpose(2, 3).double() for t in qkv.chunk(3, dim=1))
    mask = ((torch.arange(T)[None, :] << shift) < lens[:, None]).double()             # (B, T) 0/1, decoder.py:396-407
    ref = F.scaled_dot_product_attention(q, k, v, attn_mask=mask[:, None, None, :])     # float mask: additive
    ref = ref.transpose(2, 3).reshape(B, H * 64, T)
    out = torch.empty(B, H * 64, T, device="cuda")
    qd, ld = qkv.cuda(), lens.cuda()
    ctx.check(_lib.lib().ev_test_attention(ctx.handle, _lib.ptr(qd), _lib.ptr(ld), B, T, H, shift, _lib.PREC[prec],
                                           _lib.ptr(out), _lib.stream_ptr()), "ev_test_attention")
    assert rel_l2(out.cpu(), ref) < (2e-6 if prec == "fp32" else 6e-3)    # bf16: P and the output are rounded to bf16


@pytest.mark.parametrize("impl", [0, 1])
@pytest.mark.parametrize("case", [(3, 177, 2), (2, 64, 2), (1, 1, 2), (2, 129, 1), (2, 300, 2), (1, 384, 2), (2, 65, 2), (2, 700, 2)])
def test_encoder_attention_matches_reference_semantics(ctx, case, impl):
    """text_encoder.py:223-246 in float64: RoPE on the first half of each 128-wide head, scores / sqrt(128), -1e4 where the
    query OR the key is padded, softmax over all Tx keys.  impl 0 = fp32 CUDA cores, 1 = tcgen05 with 3xFP16 split operands:
    both must be fp32-accurate (durations downstream are compared bit for bit)."""
    B, T, H = case
    g = torch.Generator().manual_seed(7 * B + T)
    qkv = torch.randn(B, T, 3 * H * 128, generator=g) * 1.7
    lens = torch.randint(1, T + 1, (B,), generator=g)
    lens[0] = T
    if B > 2:
        lens[1] = 1
    q, k, v = (z.reshape(B, T, H, 128).transpose(1, 2).double() for z in qkv.chunk(3, dim=2))      # b h t c
    q, k = mo._rope(q, 64).double(), mo._rope(k, 64).double()
    m = (torch.arange(T)[None, :] < lens[:, None]).double()
    attn_mask = (m[:, None, :, None] * m[:, None, None, :])
    scores = (q @ k.transpose(-2, -1)) / math.sqrt(128)
    scores = scores.masked_fill(attn_mask == 0, -1e4)
    ref = (torch.softmax(scores, dim=-1) @ v).transpose(1, 2).reshape(B, T, H * 128)
    out = torch.full((B, T, H * 128), float("nan"), device="cuda")
    qd, ld = qkv.cuda(), lens.cuda()
    ctx.check(_lib.lib().ev_test_encoder_attention(ctx.handle, _lib.ptr(qd), _lib.ptr(ld), B, T, H, impl, _lib.ptr(out), 0, None,
                                                   _lib.stream_ptr()), "ev_test_encoder_attention")
    valid = m.bool()[:, :, None].expand_as(ref)
    err = rel_l2(out.cpu()[valid], ref[valid])
    assert err < 2e-6, err
    # padded query rows: every in-range key gets -1e4, i.e. a uniform softmax -> the mean of v
    assert torch.isfinite(out).all()
    assert rel_l2(out.cpu(), ref) < 2e-6


@pytest.mark.parametrize("case", [(2, 200, 0), (3, 129, 1), (1, 1, 0), (32, 334, 1), (5, 668, 0), (40, 1100, 0)])   # the last: several tiles per CTA
def test_fused_feed_forward_block_matches_torch(ctx, case):
    """ff_tc.cu: LayerNorm -> Linear -> SnakeBeta -> Linear -> + residual -> * mask in one kernel, against fp64 torch."""
    B, T, shift = case
    D, inner = 256, 1024
    g = torch.Generator().manual_seed(1000 + T)
    x = torch.randn(B, T, D, generator=g) * 1.5 + 0.3
    ln_g, ln_b = 1 + 0.1 * torch.randn(D, generator=g), 0.1 * torch.randn(D, generator=g)
    w1, b1 = torch.randn(inner, D, generator=g) / 16, 0.1 * torch.randn(inner, generator=g)
    w2, b2 = torch.randn(D, inner, generator=g) / 32, 0.1 * torch.randn(D, generator=g)
    sa, sb = torch.exp(0.3 * torch.randn(inner, generator=g)), 1 / (torch.exp(0.3 * torch.randn(inner, generator=g)) + 1e-9)
    lens = torch.randint(1, (T << shift) + 1, (B,), generator=g)
    lens[0] = T << shift
    if B > 2:
        lens[1], lens[2] = 1, max(1, (T << shift) // 3)     # whole tiles of padding: zero-filled, not computed
    dev = [t.cuda().contiguous() for t in (x, ln_g, ln_b, w1, b1, sa, sb, w2, b2)]
    lens_d = lens.cuda()
    out = torch.empty(B, T, D, device="cuda")
    ctx.check(_lib.lib().ev_test_ff_block(ctx.handle, *[_lib.ptr(t) for t in dev], _lib.ptr(lens_d), B, T, inner, shift,
                                          _lib.ptr(out), 0, None, _lib.stream_ptr()), "ev_test_ff_block")
    xd = x.double()
    n = F.layer_norm(xd, (D,), ln_g.double(), ln_b.double(), 1e-5)
    h = n @ w1.double().T + b1.double()
    h = h + sb.double() * torch.sin(h * sa.double()) ** 2
    y = xd + h @ w2.double().T + b2.double()
    mask = ((torch.arange(T)[None, :] << shift) < lens[:, None]).double()[:, :, None]
    ref = y * mask
    assert torch.isfinite(out).all()
    assert rel_l2(out.cpu(), ref) < 6e-3
    assert float(out.cpu()[mask.expand_as(ref) == 0].abs().sum()) == 0.0


@pytest.mark.parametrize("case", [(2, 200, 0), (3, 129, 1), (1, 1, 0), (32, 334, 1), (5, 668, 0), (40, 300, 0), (40, 1100, 0), (64, 334, 1)])
def test_fused_transformer_tail_matches_torch(ctx, case):
    """ff_tc.cu, attention tail mode: x = xr + att Wo^T + bo (out-projection + residual, transformer.py:283-294) -> LayerNorm ->
    Linear -> SnakeBeta -> Linear -> + x -> * mask in ONE kernel; the residual stream is read channel-first.  Against fp64 torch
    on the same bf16-rounded attention operand."""
    B, T, shift = case
    D, inner, A = 256, 1024, 128
    g = torch.Generator().manual_seed(2000 + T)
    xr = torch.randn(B, D, T, generator=g) * 1.5 + 0.3                     # channel-first
    att = torch.randn(B, T, A, generator=g)
    wo, bo = torch.randn(D, A, generator=g) / 11, 0.1 * torch.randn(D, generator=g)
    ln_g, ln_b = 1 + 0.1 * torch.randn(D, generator=g), 0.1 * torch.randn(D, generator=g)
    w1, b1 = torch.randn(inner, D, generator=g) / 16, 0.1 * torch.randn(inner, generator=g)
    w2, b2 = torch.randn(D, inner, generator=g) / 32, 0.1 * torch.randn(D, generator=g)
    sa, sb = torch.exp(0.3 * torch.randn(inner, generator=g)), 1 / (torch.exp(0.3 * torch.randn(inner, generator=g)) + 1e-9)
    lens = torch.randint(1, (T << shift) + 1, (B,), generator=g)
    lens[0] = T << shift
    if B > 2:
        lens[1], lens[2] = 1, max(1, (T << shift) // 3)     # whole tiles of padding: zero-filled, not computed
    dev = [t.cuda().contiguous() for t in (xr, att, wo, bo, ln_g, ln_b, w1, b1, sa, sb, w2, b2)]
    lens_d = lens.cuda()
    out = torch.empty(B, T, D, device="cuda")
    ctx.check(_lib.lib().ev_test_tf_tail(ctx.handle, *[_lib.ptr(t) for t in dev], _lib.ptr(lens_d), B, T, inner, shift,
                                         _lib.ptr(out), 0, None, _lib.stream_ptr()), "ev_test_tf_tail")
    bf = lambda t: t.bfloat16().double()
    xd = xr.double().transpose(1, 2) + bf(att) @ bf(wo).T + bo.double()
    n = F.layer_norm(xd, (D,), ln_g.double(), ln_b.double(), 1e-5)
    h = n @ w1.double().T + b1.double()
    h = h + sb.double() * torch.sin(h * sa.double()) ** 2
    y = xd + h @ w2.double().T + b2.double()
    mask = ((torch.arange(T)[None, :] << shift) < lens[:, None]).double()[:, :, None]
    ref = y * mask
    assert torch.isfinite(out).all()
    assert rel_l2(out.cpu(), ref) < 6e-3
    assert float(out.cpu()[mask.expand_as(ref) == 0].abs().sum()) == 0.0


def _mish(x):
    return x * torch.tanh(F.softplus(x))


# B, T, C_in, len_shift, full      (one tile per CTA, 126 owned frames per m-block + 2 halo rows; B * ceil(T / 126) <= 148 -> one
# m-block per CTA, else two; (50, 300) ends in a tile whose second m-block is empty, 126 / 127 / 252 sit on tile boundaries)
RESNET_CASES = [(2, 200, 256, 0, 1), (3, 129, 224, 1, 1), (1, 1, 256, 0, 1), (32, 334, 512, 1, 1), (32, 668, 224, 0, 1),
                (5, 700, 512, 0, 1), (32, 668, 256, 0, 0), (2, 130, 256, 0, 0), (40, 500, 256, 0, 1), (50, 300, 256, 0, 1),
                (3, 126, 256, 0, 1), (3, 127, 256, 0, 1), (2, 252, 256, 1, 1), (50, 254, 512, 0, 1), (50, 255, 256, 0, 0),
                (200, 130, 256, 0, 1), (3, 1000, 256, 0, 1), (2, 1500, 224, 0, 1), (20, 1008, 256, 0, 0),   # several waves; clusters of 8 and 6
                (32, 1100, 512, 0, 1), (32, 334, 256, 1, 2), (32, 668, 512, 0, 2), (3, 129, 224, 1, 2)]     # full = 2: channel-first stream


@pytest.mark.parametrize("case", RESNET_CASES)
def test_fused_resnet_block_matches_torch(ctx, case):
    """resnet_tc.cu: conv3 -> GroupNorm(8) over the padded extent -> Mish -> mask -> + temb -> mask -> conv3 -> GroupNorm -> Mish ->
    mask -> + res_conv(x) -> LayerNorm, one launch, the tiles of an item exchanging GroupNorm sums inside a thread-block cluster, against fp64 torch on the same
    bf16-rounded operands (decoder.py:32-61, transformer.py:262)."""
    B, T, Cin, shift, full = case
    D = 256
    g = torch.Generator().manual_seed(B * 7919 + T * 31 + Cin)
    x = torch.randn(B, Cin, T, generator=g) * 1.2
    lens = torch.randint(1, (T << shift) + 1, (B,), generator=g)
    lens[0] = T << shift
    if B > 2:
        lens[1], lens[2] = 1, max(1, (T << shift) // 3)
    rnd = lambda *s, scale=1.0: torch.randn(*s, generator=g) * scale
    w = {"conv1.weight": rnd(D, Cin, 3, scale=(2.0 / (3 * Cin)) ** 0.5), "conv1.bias": rnd(D, scale=0.1),
         "gn1.weight": 1 + rnd(D, scale=0.1), "gn1.bias": rnd(D, scale=0.1)}
    if full:
        w.update({"temb": rnd(D, scale=0.5), "conv2.weight": rnd(D, D, 3, scale=(2.0 / (3 * D)) ** 0.5), "conv2.bias": rnd(D, scale=0.1),
                  "gn2.weight": 1 + rnd(D, scale=0.1), "gn2.bias": rnd(D, scale=0.1), "res.weight": rnd(D, Cin, 1, scale=Cin ** -0.5),
                  "res.bias": rnd(D, scale=0.1), "ln.weight": 1 + rnd(D, scale=0.1), "ln.bias": rnd(D, scale=0.1)})
    arr, keep = _lib.tensor_list(w, ctx.device)
    xd, ld = x.cuda(), lens.cuda()
    out_a, out_xr, out_n = (torch.empty(B, T, D, device="cuda") for _ in range(3))
    ctx.check(_lib.lib().ev_test_resnet_block(ctx.handle, arr, len(keep), _lib.ptr(xd), _lib.ptr(ld), B, T, Cin, shift, full,
                                              _lib.ptr(out_a), _lib.ptr(out_xr), _lib.ptr(out_n), 0, None, _lib.stream_ptr()),
              "ev_test_resnet_block")
    # reference on the same bf16-rounded operands, fp64
    bf = lambda t: t.bfloat16().double()
    m = ((torch.arange(T)[None, :] << shift) < lens[:, None]).double()[:, None, :]          # (B, 1, T)
    xm = bf(x * m.float())
    h = F.conv1d(xm, bf(w["conv1.weight"]), w["conv1.bias"].double(), padding=1)
    h = _mish(F.group_norm(h, 8, w["gn1.weight"].double(), w["gn1.bias"].double(), 1e-5)) * m
    if not full:
        assert torch.isfinite(out_a).all()
        assert rel_l2(out_a.cpu(), h.transpose(1, 2)) < 4e-3
        return
    a_ref = (h + w["temb"].double()[None, :, None]) * m
    h2 = F.conv1d(bf(a_ref.float()), bf(w["conv2.weight"]), w["conv2.bias"].double(), padding=1)
    h2 = _mish(F.group_norm(h2, 8, w["gn2.weight"].double(), w["gn2.bias"].double(), 1e-5)) * m
    xr_ref = h2 + F.conv1d(xm, bf(w["res.weight"]), w["res.bias"].double())
    n_ref = F.layer_norm(xr_ref.transpose(1, 2), (D,), w["ln.weight"].double(), w["ln.bias"].double(), 1e-5)
    for t in (out_a, out_xr, out_n):
        assert torch.isfinite(t).all()
    assert rel_l2(out_a.cpu(), a_ref.transpose(1, 2)) < 4e-3             # bf16 output
    assert rel_l2(out_xr.cpu(), xr_ref.transpose(1, 2)) < 4e-3           # conv2 consumes the bf16-rounded `a` (1-ulp flips vs the reference's)
    assert rel_l2(out_n.cpu(), n_ref) < 6e-3
    assert float(out_a.cpu()[(m.transpose(1, 2) == 0).expand_as(out_a)].abs().sum()) == 0.0


def test_length_sum_follows_aten_cpu_order(ctx):
    rng = np.random.default_rng(0)
    for n in (1, 3, 7, 8, 9, 17, 151, 333, 513, 1100, 2100):
        for ls in (0.8, 0.9, 1.0, 1.1, 1.2):
            w = torch.from_numpy((np.ceil(np.exp(rng.normal(0.9, 0.6, size=(32, n)))) * ls).astype(np.float32))
            out = torch.empty(32, device="cuda")
            wd = w.cuda()
            ctx.check(_lib.lib().ev_test_row_sum(ctx.handle, _lib.ptr(wd), 32, n, _lib.ptr(out), _lib.stream_ptr()), "row_sum")
            assert torch.equal(out.cpu(), torch.sum(w.view(32, 1, n), [1, 2])), (n, ls)      # bit-exact


# ------------------------------------------------------------------------------------------------ synthesise
def _near_tie_tokens(ref):
    w = torch.exp(ref["logw"]) * ref["x_mask"]
    return ((w - torch.round(w)).abs() < 1e-4) & (ref["x_mask"] > 0)


def _check_synthesise(out, ref, tol):
    # durations are bit-exact.  (ceil(exp(logw)) could only flip where the oracle's own duration sits within 1e-4 of an
    # integer; none of the fixtures / seeds used here holds such a token -- a mismatch is a failure, not a skip.)
    assert torch.equal(out["w_ceil"].cpu(), ref["w_ceil"]), \
        f"duration mismatch ({int((out['w_ceil'].cpu() != ref['w_ceil']).sum())} tokens, {int(_near_tie_tokens(ref).sum())} near-ties in the batch)"
    assert out["mel_lengths"].cpu().tolist() == ref["mel_lengths"].tolist()
    assert torch.equal(out["attn"].cpu(), ref["attn"])
    assert rel_l2(out["encoder_outputs"].cpu(), ref["encoder_outputs"]) < 2e-5
    # mu_y = attn^T mu_x is implemented as a gather: it must reproduce the GPU's own mu_x exactly
    a = out["attn"][:, 0].cpu()                                                          # (B, Tx, T_pad)
    gathered = torch.matmul(a.transpose(1, 2), out["mu_x"].cpu().transpose(1, 2)).transpose(1, 2)
    ymax = int(out["mel_lengths"].max())
    assert torch.equal(out["encoder_outputs"].cpu(), gathered[:, :, :ymax])
    assert rel_l2(out["decoder_outputs"].cpu(), ref["decoder_outputs"]) < tol
    assert rel_l2(out["mel"].cpu(), ref["mel"]) < tol
    assert rel_l2(out["decoder_outputs_full"].cpu(), ref["decoder_outputs_full"]) < tol  # padded frames too (H1)


@pytest.mark.parametrize("prec", ["fp32", "bf16"])
@pytest.mark.parametrize("name", golden_io.MATCHA)
def test_synthesise_matches_oracle_and_reference_fixture(matcha, matcha_sd, name, prec):
    g = golden_io.load(name)
    inp = golden_io.matcha_inputs(g)
    ref = mo.synthesise(matcha_sd, VCTK, **inp)
    out = matcha.synthesise(inp["x"], inp["x_lengths"], inp["n_timesteps"], inp["temperature"], inp["spks"],
                            inp["length_scale"], z=inp["z"], dtype=prec)
    assert rel_l2(out["logw"].cpu(), ref["logw"]) < 2e-5
    assert rel_l2(out["mu_x"].cpu(), ref["mu_x"]) < 2e-5
    _check_synthesise(out, ref, TOL[prec])
    # and against what the reference itself wrote
    assert out["mel_lengths"].cpu().tolist() == g["mel_lengths"].tolist()
    assert torch.equal(out["attn"][:, 0].cpu(), g["attn"])
    assert rel_l2(out["mel"].cpu(), torch.from_numpy(g["mel"])) < TOL[prec]


def test_bf16_mel_cepstral_distortion_is_reported_and_small(matcha, matcha_sd, capsys):
    """BASELINE north star: bf16 mel within rel-L2 1e-2 of the reference path 'with mel-cepstral distortion reported'.  MCD (13
    coefficients, valid frames) of the tensor-core path against the fp32 oracle on the same inputs and prior noise; the fp32
    CUDA path is reported alongside.  Typical TTS systems sit at 4-8 dB from their ground truth; numerical noise must be
    orders of magnitude below that."""
    x, xl, spk = synthetic.phoneme_batch(3, 12, 30, seed=11)
    probe = mo.synthesise(matcha_sd, VCTK, x, xl, 1, 0.667, spk, 0.8)
    z = synthetic.prior_noise(3, 80, probe["t_pad"], seed=12)
    ref = mo.synthesise(matcha_sd, VCTK, x, xl, 10, 0.667, spk, 0.8, z=z)
    res = {}
    for prec in ("fp32", "bf16"):
        out = matcha.synthesise(x, xl, 10, 0.667, spk, 0.8, z=z, dtype=prec)
        res[prec] = (mel_cepstral_distortion(out["mel"], ref["mel"], ref["mel_lengths"]), rel_l2(out["mel"].cpu(), ref["mel"]))
    with capsys.disabled():
        print(f"\nMCD vs oracle: fp32 {res['fp32'][0]:.2e} dB (rel-L2 {res['fp32'][1]:.1e}), bf16 {res['bf16'][0]:.2e} dB (rel-L2 {res['bf16'][1]:.1e})")
    assert res["fp32"][0] < 1e-3 and res["bf16"][0] < 1.0      # measured on B200: 5.7e-5 dB and 0.34 dB
    assert res["bf16"][1] < TOL["bf16"]


@pytest.mark.parametrize("ls", [0.8, 0.9, 1.0, 1.1, 1.2])
def test_durations_and_alignment_bit_exact_over_length_scales(matcha, matcha_sd, ls):
    # A batch whose oracle durations hold a near-tie token (|w - round(w)| < 1e-4, SURVEY H2a: the one place where a last-ulp
    # difference of exp() may legitimately flip a ceil) is replaced by the next seed instead of being skipped.
    for seed in range(int(ls * 100), int(ls * 100) + 5000, 1000):
        x, xl, spk = synthetic.phoneme_batch(8, 2, 60, seed=seed)
        ref = mo.synthesise(matcha_sd, VCTK, x, xl, 1, 0.667, spk, ls)
        if not bool(_near_tie_tokens(ref).any()):
            break
    else:
        raise AssertionError("no near-tie-free batch among 5 seeds")
    out = matcha.synthesise(x, xl, 1, 0.667, spk, ls, z=torch.zeros(8, 80, ref["t_pad"]), dtype="fp32")
    assert torch.equal(out["w_ceil"].cpu(), ref["w_ceil"])
    assert out["mel_lengths"].cpu().tolist() == ref["mel_lengths"].tolist()
    assert out["t_pad"] == ref["t_pad"]
    assert torch.equal(out["attn"].cpu(), ref["attn"])
    attn = out["attn"][:, 0].cpu()
    assert bool(((attn == 0) | (attn == 1)).all())
    assert torch.equal(attn.sum(1), ref["y_mask"][:, 0])                               # one token per valid frame


def test_padded_frames_carry_scaled_noise(matcha, matcha_sd):
    """the estimator output is masked, so padded frames of decoder_outputs equal z*temperature exactly."""
    x, xl, spk = synthetic.phoneme_batch(2, 5, 25, seed=21)
    probe = mo.synthesise(matcha_sd, VCTK, x, xl, 1, 0.5, spk, 1.0)
    z = synthetic.prior_noise(2, 80, probe["t_pad"], seed=22)
    out = matcha.synthesise(x, xl, 3, 0.5, spk, 1.0, z=z, dtype="bf16")
    full, lens = out["decoder_outputs_full"].cpu(), out["mel_lengths"].cpu()
    for b in range(2):
        assert torch.equal(full[b, :, lens[b]:], (z * 0.5)[b, :, lens[b]:])


def test_single_speaker_edge_cases(matcha, matcha_sd):
    # shortest possible text (one blank), B=1; and a ragged batch where one item is a single token
    x = torch.zeros(1, 1, dtype=torch.long)
    xl = torch.tensor([1])
    spk = torch.tensor([12])
    ref = mo.synthesise(matcha_sd, VCTK, x, xl, 2, 0.667, spk, 1.0, z=torch.zeros(1, 80, 4))
    out = matcha.synthesise(x, xl, 2, 0.667, spk, 1.0, z=torch.zeros(1, 80, ref["t_pad"]), dtype="fp32")
    assert out["mel_lengths"].cpu().tolist() == ref["mel_lengths"].tolist()
    assert rel_l2(out["mel"].cpu(), ref["mel"]) < 1e-4
    with pytest.raises(ValueError):
        matcha.synthesise(x, xl, 2, spks=None)
    with pytest.raises(ValueError):
        matcha.synthesise(x, xl, 2, spks=spk, z=torch.zeros(1, 80, 3))


# ------------------------------------------------------------------------------------------------ vocoder
@pytest.mark.parametrize("prec", ["fp32", "bf16"])
@pytest.mark.parametrize("name", list(golden_io.HIFIGAN))
def test_vocoder_matches_oracle_and_reference_fixture(vocoders, name, prec):
    gen, sd = vocoders[name]
    g = golden_io.load(name)
    b, frames, seed = (int(v) for v in g["meta"])
    mel = synthetic.synthetic_mel(b, frames, seed=seed)
    wav = gen(mel, dtype=prec)
    assert wav.shape == (b, 1, frames * 256)
    assert rel_l2(wav.cpu(), ho.generator(sd, HIFIGAN_V1, mel)) < TOL[prec]
    assert rel_l2(wav.cpu(), torch.from_numpy(g["wav"])) < TOL[prec]
    assert float(wav.abs().max()) <= 1.0


@pytest.mark.parametrize("name", list(golden_io.HIFIGAN))
def test_denoiser_matches_oracle_and_reference_fixture(vocoders, name):
    """hifigan/denoiser.py: bias spectrum of vocoder(zeros) and the STFT -> subtract -> ISTFT path, against the oracle
    and against what the reference itself produced (tests/golden)."""
    gen, sd = vocoders[name]
    g = golden_io.load(name)
    b, frames, seed = (int(v) for v in g["meta"])
    gen.precision = "fp32"                              # the bias spectrum is computed with the generator's own precision
    try:
        den = ev.Denoiser(gen, mode="zeros")
        wav = gen(synthetic.synthetic_mel(b, frames, seed=seed)).clamp(-1, 1)
    finally:
        gen.precision = "bf16"
    bias_ref = ho.denoiser_bias(sd, HIFIGAN_V1)
    assert den.bias_spec.shape == bias_ref.shape == (1, 513, 1)
    assert rel_l2(den.bias_spec.cpu(), bias_ref) < 1e-4
    assert rel_l2(den.bias_spec.cpu(), torch.from_numpy(g["bias_spec"])) < 1e-4
    for strength in (0.00025, 0.0005, 0.1):            # app default (feel_me.py:185), class default, a strong setting
        out = den(wav.squeeze(1), strength=strength)
        ref = ho.denoise(wav.cpu().squeeze(1), bias_ref, strength)
        assert out.shape == ref.shape
        assert rel_l2(out.cpu(), ref) < 1e-4
    out = den(wav.squeeze(1), strength=0.00025)
    assert rel_l2(out.cpu(), torch.from_numpy(g["denoised"])) < 2e-4
    one = den(wav[0, 0], strength=0.00025)             # 1-D input, as to_waveform passes it after squeeze()
    assert one.dim() == 1 and torch.equal(one, out[0])
    with pytest.raises(Exception):
        ev.Denoiser(gen, mode="normal")


def test_vocoder_accepts_weight_norm_checkpoint_form():
    wn = synthetic.hifigan_state_dict(HIFIGAN_V1, seed=7, gain=0.5, weight_norm=True)
    gen = ev.Generator(HIFIGAN_V1)
    gen.load_state_dict(wn)
    gen.eval()
    gen.remove_weight_norm()
    mel = synthetic.synthetic_mel(1, 19, seed=8)
    ref = ho.generator(ho.fold_weight_norm(wn), HIFIGAN_V1, mel)
    assert rel_l2(gen(mel, dtype="fp32").cpu(), ref) < 1e-4


def test_vocoder_is_linear_in_batch_and_time_tiling_free():
    """Size-independent property at a larger size: each utterance's waveform does not depend on its batch-mates."""
    sd = synthetic.hifigan_state_dict(HIFIGAN_V1, seed=4321, gain=1.0)
    gen = ev.Generator(HIFIGAN_V1)
    gen.load_state_dict(sd)
    gen.remove_weight_norm()
    mel = synthetic.synthetic_mel(4, 200, seed=31)
    full = gen(mel, dtype="bf16")
    solo = gen(mel[2:3], dtype="bf16")
    assert torch.equal(full[2:3], solo)


@pytest.mark.parametrize("prec", ["bf16", "fp32"])
def test_ragged_vocoder_equals_dense_on_valid_samples_and_is_zero_beyond(vocoders, prec):
    """ev_vocode_ragged: item b's waveform is BIT-identical to the dense generator on [: len_b * 256] (what the reference's
    batched caller keeps, cli.py:307-311) and exactly zero beyond; lengths 0, 1, T and > T, eager and graph replay, and
    replays of one graph with different lengths."""
    gen, sd = vocoders["hifigan_gain1"]
    B, T = 6, 300
    mel = synthetic.synthetic_mel(B, T, seed=77)
    dense = gen(mel, dtype=prec)
    for lens in ([300, 0, 1, 137, 299, 512], [17, 300, 300, 64, 2, 250], [100] * 6):
        for _ in range(3):                               # eager, capture, replay
            wav = gen(mel, dtype=prec, lengths=torch.tensor(lens))
            assert wav.shape == dense.shape
            for b, n in enumerate(lens):
                n = min(n, T) * 256
                assert torch.equal(wav[b, :, :n], dense[b, :, :n]), (lens, b)
                assert float(wav[b, :, n:].abs().max() if n < T * 256 else 0.0) == 0.0
    # the valid part also matches the oracle's (reference's) cropped waveform
    ref = ho.generator(sd, HIFIGAN_V1, mel)
    wav = gen(mel, dtype=prec, lengths=torch.tensor([137] * B))
    assert rel_l2(wav[:, :, : 137 * 256].cpu(), ref[:, :, : 137 * 256]) < TOL[prec]
    # a short batch (layer-by-layer ResBlocks below 1024 samples per stage) and a lone long item
    mel_s = synthetic.synthetic_mel(3, 20, seed=78)
    d_s, r_s = gen(mel_s, dtype=prec), gen(mel_s, dtype=prec, lengths=[20, 3, 11])
    for b, n in enumerate([20, 3, 11]):
        assert torch.equal(r_s[b, :, : n * 256], d_s[b, :, : n * 256]) and float(r_s[b, :, n * 256:].abs().sum()) == 0.0


def test_ragged_vocoder_long_segments(vocoders):
    """BASELINE config 5 shape (tens of seconds per segment): thousands of tiles per item in the last stages."""
    gen, _ = vocoders["hifigan_gain1"]
    mel = synthetic.synthetic_mel(2, 2000, seed=79)
    dense = gen(mel)
    rag = gen(mel, lengths=[2000, 700])
    assert torch.equal(rag[0], dense[0])
    assert torch.equal(rag[1, :, : 700 * 256], dense[1, :, : 700 * 256]) and float(rag[1, :, 700 * 256:].abs().sum()) == 0.0


def test_ragged_vocoder_full_size_batch(matcha, matcha_sd, vocoders):
    """BASELINE config 2 shape (32 x ~600 frames, mixed lengths) through synthesise -> ragged vocoder: valid samples equal
    the dense path bit for bit, the corpus driver's cropped waveforms are unchanged by `ragged`."""
    gen, _ = vocoders["hifigan_gain1"]
    x, xl, spks = synthetic.phoneme_batch(32, 60, 90, seed=2000)
    out = matcha.synthesise(x, xl, 2, 0.667, spks, 0.8)
    dense = gen(out["mel"])
    rag = gen(out["mel"], lengths=out["mel_lengths"])
    ml = out["mel_lengths"].cpu().tolist()
    assert min(ml) < max(ml)
    for b, n in enumerate(ml):
        assert torch.equal(rag[b, :, : n * 256], dense[b, :, : n * 256])
        assert float(rag[b, :, n * 256:].abs().sum()) == 0.0
    utts = [(x[i, : int(xl[i])].tolist(), int(spks[i])) for i in range(8)]
    zs = {}

    def z_fn(mb, model, x_, xl_, spks_):
        probe = model.synthesise(x_, xl_, 1, 0.667, spks_, 1.0)
        return zs.setdefault(tuple(mb.items), synthetic.prior_noise(len(mb.items), 80, probe["t_pad"], seed=3))

    r0, _ = ev.synthesise_corpus(matcha, gen, utts, batch_size=4, n_timesteps=2, z_fn=z_fn, ragged=False)
    r1, _ = ev.synthesise_corpus(matcha, gen, utts, batch_size=4, n_timesteps=2, z_fn=z_fn, ragged=True)
    for i in r0:
        assert torch.equal(r0[i]["waveform"], r1[i]["waveform"])


def test_corpus_lanes_do_not_change_results(matcha, matcha_sd, vocoders):
    """Several micro-batches in flight (ev.Lanes: model / vocoder replicas on their own streams) give the waveforms of the
    one-at-a-time loop bit for bit, with and without the denoiser; stats count every utterance once."""
    gen, _ = vocoders["hifigan_gain1"]
    x, xl, spks = synthetic.phoneme_batch(10, 5, 30, seed=21)
    utts = [(x[i, : int(xl[i])].tolist(), int(spks[i])) for i in range(10)]
    zs = {}

    def z_fn(mb, model, x_, xl_, spks_):
        probe = model.synthesise(x_, xl_, 1, 0.667, spks_, 1.0)
        return zs.setdefault(tuple(mb.items), synthetic.prior_noise(len(mb.items), 80, probe["t_pad"], seed=5))

    den = ev.Denoiser(gen)
    for denoiser in (None, den):
        r1, s1 = ev.synthesise_corpus(matcha, gen, utts, batch_size=2, n_timesteps=2, z_fn=z_fn, denoiser=denoiser)
        r3, s3 = ev.synthesise_corpus(matcha, gen, utts, batch_size=2, n_timesteps=2, z_fn=z_fn, denoiser=denoiser, in_flight=3)
        assert sorted(r1) == sorted(r3) == list(range(10))
        assert s1.utterances == s3.utterances == 10 and s3.extra["in_flight"] == 3 and s3.seconds > 0
        for i in r1:
            assert r1[i]["mel_length"] == r3[i]["mel_length"]
            assert torch.equal(r1[i]["waveform"], r3[i]["waveform"])
    assert len(ev.lanes_for(matcha, gen, 3)) == 3 and ev.lanes_for(matcha, gen, 3) is ev.lanes_for(matcha, gen, 3)


def test_end_to_end_emoji_text_to_waveform(matcha, matcha_sd, vocoders):
    gen, hsd = vocoders["hifigan_gain1"]
    text, spk = ev.emoji_to_spk("that is wonderful \U0001F60D")
    assert spk == 107
    ids = torch.tensor([ev.intersperse([(ord(c) % 150) + 1 for c in text.strip()])])
    xl = torch.tensor([ids.shape[1]])
    spks = torch.tensor([spk])
    probe = mo.synthesise(matcha_sd, VCTK, ids, xl, 1, 0.667, spks, 0.8)
    z = synthetic.prior_noise(1, 80, probe["t_pad"], seed=41)
    ref = mo.synthesise(matcha_sd, VCTK, ids, xl, 10, 0.667, spks, 0.8, z=z)
    out = matcha.synthesise(ids, xl, 10, 0.667, spks, 0.8, z=z, dtype="fp32")
    assert out["mel_lengths"].cpu().tolist() == ref["mel_lengths"].tolist()
    ref_wav = ho.to_waveform(hsd, HIFIGAN_V1, ref["mel"])
    for prec in ("fp32", "bf16"):                      # to_waveform uses the generator's own precision setting
        gen.precision = prec
        try:
            wav = ev.to_waveform(out["mel"], gen)
        finally:
            gen.precision = "bf16"
        assert wav.shape == ref_wav.shape
        assert rel_l2(wav, ref_wav) < TOL[prec]


def test_corpus_driver_matches_oracle_on_identical_microbatches(matcha_sd, vocoders):
    """the micro-batch list is part of the input -- the oracle is run on the same micro-batches, with the
    same prior noise; every utterance's cropped waveform (cli.py:308-309) must match.  Two "ranks" cover the list."""
    gen, hsd = vocoders["hifigan_gain1"]
    model = ev.MatchaTTS(**VCTK.constructor_kwargs(), precision="fp32")
    model.load_state_dict(matcha_sd)
    g = torch.Generator().manual_seed(77)
    utts = []
    for i, p in enumerate((9, 4, 14, 6, 11)):
        ids = ev.intersperse(torch.randint(1, 178, (p,), generator=g).tolist())
        utts.append((ids, list(ev.EMOJI_MAPPING_FEMALE.values())[i]))
    ref_wavs, noise = {}, {}

    def z_fn(mb, model_, x, xl, spks):
        probe = mo.synthesise(matcha_sd, VCTK, x, xl, 1, 0.667, spks, 0.8)
        noise[mb.index] = synthetic.prior_noise(x.shape[0], 80, probe["t_pad"], seed=100 + mb.index)
        ref = mo.synthesise(matcha_sd, VCTK, x, xl, 2, 0.667, spks, 0.8, z=noise[mb.index])
        wav = ho.generator(hsd, HIFIGAN_V1, ref["mel"]).clamp(-1, 1)
        for j, i in enumerate(mb.items):
            ref_wavs[i] = wav[j, 0, : int(ref["mel_lengths"][j]) * 256]
        return noise[mb.index]

    gen.precision = "fp32"
    try:
        got = {}
        for rank in range(2):
            res, stats = ev.synthesise_corpus(model, gen, utts, batch_size=2, n_timesteps=2, temperature=0.667, length_scale=0.8,
                                              rank=rank, world_size=2, z_fn=z_fn)
            assert stats.utterances == len(res) and stats.seconds > 0
            assert not set(res) & set(got)
            got.update(res)
    finally:
        gen.precision = "bf16"
    assert sorted(got) == list(range(5))
    for i in range(5):
        assert got[i]["waveform"].shape == ref_wavs[i].shape == (got[i]["mel_length"] * 256,)
        assert rel_l2(got[i]["waveform"], ref_wavs[i]) < 2e-4


# ------------------------------------------------------------------------------------------------ full size (config 2)
def test_full_size_batch_properties_and_anchor_items(matcha, matcha_sd, vocoders):
    """BASELINE configs[1] at its full size (32 utterances of ~5 s, n_timesteps = 10), through size-independent properties:
      * the two items the oracle can afford -- the longest utterance (it fixes T_pad, so the oracle's own 2-item batch pads
        exactly like the 32-item one; SURVEY H1: items only interact through the padded extent) and one more -- match the
        oracle within the fp32 tolerance; their durations / alignment are bit-exact;
      * every alignment is a monotonic 0/1 path with one token per valid frame; padded frames carry z * temperature;
      * bf16 (tensor-core) and fp32 (CUDA-core) modes agree within the bf16 tolerance on the whole batch, mel and waveform;
      * the CUDA-graph replay (third call with the same shape) is bit-identical to the eager call."""
    gen, hsd = vocoders["hifigan_gain1"]
    x, xl, spk = synthetic.phoneme_batch(32, 60, 90, seed=2000)
    probe = matcha.synthesise(x, xl, 1, 0.667, spk, 0.8, dtype="fp32")
    t_pad, lens = probe["t_pad"], probe["mel_lengths"].cpu()
    z = synthetic.prior_noise(32, 80, t_pad, seed=7)
    out32 = matcha.synthesise(x, xl, 10, 0.667, spk, 0.8, z=z, dtype="fp32")
    assert torch.equal(out32["mel_lengths"].cpu(), lens)
    # anchors: longest item + item 0 (or 1) in one oracle batch -> same T_pad
    i_long = int(lens.argmax())
    i_other = 0 if i_long != 0 else 1
    sel = torch.tensor([i_long, i_other])
    tx = int(xl[sel].max())
    ref = mo.synthesise(matcha_sd, VCTK, x[sel, :tx], xl[sel], 10, 0.667, spk[sel], 0.8, z=z[sel])
    assert ref["t_pad"] == t_pad
    assert ref["mel_lengths"].tolist() == lens[sel].tolist()
    attn = out32["attn"][:, 0].cpu()
    assert torch.equal(attn[sel][:, :tx], ref["attn"][:, 0])
    assert rel_l2(out32["mel"].cpu()[sel], ref["mel"]) < 1e-4
    # alignment properties on all 32 items
    assert bool(((attn == 0) | (attn == 1)).all())
    frames = torch.arange(t_pad)[None, :] < lens[:, None]
    assert torch.equal(attn.sum(1), frames.float())
    tok = attn.argmax(1)
    for b in range(32):
        tb = tok[b, : int(lens[b])]
        # non-decreasing token index per frame (with length_scale < 1 a token may receive no frame at all, so steps can exceed 1)
        assert int(tb[0]) == 0 and int(tb[-1]) <= int(xl[b]) - 1 and bool((tb[1:] - tb[:-1] >= 0).all())
    full = out32["decoder_outputs_full"].cpu()
    for b in (0, 7, 31):
        assert torch.equal(full[b, :, int(lens[b]):], (z * 0.667)[b, :, int(lens[b]):])
    # bf16 vs fp32 on the whole batch; graph replay == eager
    runs = [matcha.synthesise(x, xl, 10, 0.667, spk, 0.8, z=z, dtype="bf16") for _ in range(3)]
    assert torch.equal(runs[0]["mel"], runs[2]["mel"])
    assert rel_l2(runs[2]["mel"].cpu(), out32["mel"].cpu()) < TOL["bf16"]
    wav32 = gen(out32["mel"], dtype="fp32")
    wavs = [gen(out32["mel"], dtype="bf16") for _ in range(3)]
    assert torch.equal(wavs[0], wavs[2])
    assert wav32.shape == (32, 1, 256 * out32["mel"].shape[2])
    assert rel_l2(wavs[2].cpu(), wav32.cpu()) < TOL["bf16"]
    ref_wav = ho.generator(hsd, HIFIGAN_V1, out32["mel"].cpu()[sel[:1]])
    assert rel_l2(wav32.cpu()[sel[:1]], ref_wav) < 1e-4


# ------------------------------------------------------------------------------------------------ scheduling knobs
def test_scheduling_knobs_do_not_change_results():
    """Decoder batch lanes on forked streams, the ResBlock wavefront schedule, one vs two CTAs per SM, programmatic
    dependent launch and the ragged tile lists of the decoder only reorder or skip dead work: mel and waveform must be
    bit-identical to the default schedule."""
    import os
    import subprocess
    import sys

    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))

    def run(**env):
        e = dict(os.environ, **{k: str(v) for k, v in env.items()})
        r = subprocess.run([sys.executable, os.path.join(root, "scripts", "knob_check.py")], capture_output=True, text=True, env=e, timeout=600)
        assert r.returncode == 0, r.stderr[-2000:]
        return [l for l in r.stdout.splitlines() if l.startswith("HASH")][0]

    base = run()
    # decoder lanes and the res_conv side branch belong to the layer-by-layer ResNet path (EV_RN_FUSE=0: five launches per block
    # instead of the fused kernel, whose arithmetic order differs -- and depends on the tile plan, hence on the lane's batch
    # size): compared among themselves
    unfused = run(EV_RN_FUSE=0)
    assert run(EV_RN_FUSE=0, EV_DEC_LANES=2) == unfused
    assert run(EV_RN_FUSE=0, EV_DEC_SIDE=0) == unfused
    assert run() == base                            # the fused kernel's sums have one writer each and a fixed order: reproducible
    # HiFi-GAN ResBlocks: windows paired in 2-CTA clusters (halo rows exchanged over DSMEM) compute every row in the same order
    # as single windows -- with the same set of fused blocks (EV_RB_FUSE=2: all nine) the waveform does not change by a bit
    all_fused = run(EV_RB_FUSE=2)
    assert run(EV_RB_FUSE=2, EV_RB_CLUSTER=1) == all_fused
    assert run(EV_RB_FUSE=2, EV_RB_CLUSTER=3) == all_fused
    assert all_fused == base                        # the default policy fuses all nine as well
    # the wavefront schedule and one CTA per SM keep single windows: compared with the unpaired default (C = 128, k >= 7 layer by layer)
    single = run(EV_RB_CLUSTER=1)
    assert run(EV_RB_WAVE=1) == single
    assert run(EV_RB_WAVE=1, EV_RB_OCC2=0) == single
    assert run(EV_PDL=0) == base
    # the attention's q|k|v projection inside the fused ResNet kernel multiplies the same bf16 operands in the same K order as the
    # separate conv launch it replaces
    assert run(EV_QKV_FUSE=0) == base
    # skipping the padded rows of the decoder's masked per-row work (feed-forward tiles, attention query blocks, out-projection
    # tiles) must not change a single bit either: those rows never reach an output that survives the mask
    assert run(EV_FF_RAGGED=0) == base


def test_long_utterance_matches_oracle(matcha, matcha_sd, vocoders):
    """One ~20 s utterance (Tx = 601 tokens, ~1.8 k mel frames): many key blocks in the tensor-core attention, dozens of
    tiles per conv, RoPE far into its table; plus a 2-frame neighbour so that almost the whole batch is padding (H1)."""
    gen, hsd = vocoders["hifigan_gain1"]
    g = torch.Generator().manual_seed(91)
    ids = torch.zeros(2, 601, dtype=torch.long)
    ids[0, 1::2] = torch.randint(1, 178, (300,), generator=g)
    ids[1, :3] = torch.tensor([0, 17, 0])
    xl = torch.tensor([601, 3])
    spk = torch.tensor([18, 54])
    probe = mo.synthesise(matcha_sd, VCTK, ids, xl, 1, 0.667, spk, 1.0)
    z = synthetic.prior_noise(2, 80, probe["t_pad"], seed=92)
    ref = mo.synthesise(matcha_sd, VCTK, ids, xl, 2, 0.667, spk, 1.0, z=z)
    assert ref["t_pad"] > 1500
    for prec in ("fp32", "bf16"):
        out = matcha.synthesise(ids, xl, 2, 0.667, spk, 1.0, z=z, dtype=prec)
        assert out["mel_lengths"].cpu().tolist() == ref["mel_lengths"].tolist()
        assert torch.equal(out["attn"].cpu(), ref["attn"])
        assert rel_l2(out["mel"].cpu(), ref["mel"]) < TOL[prec]
        assert rel_l2(out["decoder_outputs_full"].cpu(), ref["decoder_outputs_full"]) < TOL[prec]
    wav = gen(out["mel"], dtype="fp32")
    ref_wav = ho.generator(hsd, HIFIGAN_V1, out["mel"].cpu())
    assert rel_l2(wav.cpu(), ref_wav) < 1e-4
