"""Architecture constants of the synthesis hot path (VCTK multi-speaker Matcha-TTS + HiFi-GAN v1).

Values restate the reference's Hydra/YAML and python config, they are read as constants only:
  Matcha-TTS/configs/model/matcha.yaml:9-12, configs/model/encoder/default.yaml:4-17,
  configs/model/decoder/default.yaml:1-7, configs/model/cfm/default.yaml:1-3,
  configs/data/vctk.yaml:11-14, matcha/hifigan/config.py:1-28.
"""
from __future__ import annotations

from dataclasses import dataclass, field, asdict


class AttrDict(dict):
    """dict with attribute access (same contract as matcha/hifigan/env.py:7-10)."""

    def __init__(self, *args, **kwargs):
        super().__init__(*args, **kwargs)
        self.__dict__ = self


@dataclass(frozen=True)
class MatchaConfig:
    n_vocab: int = 178
    n_spks: int = 109
    spk_emb_dim: int = 64
    n_feats: int = 80
    # text encoder (encoder_params)
    enc_channels: int = 192
    enc_filter_channels: int = 768
    enc_filter_channels_dp: int = 256
    enc_heads: int = 2
    enc_layers: int = 6
    enc_kernel: int = 3
    enc_prenet: bool = True
    # flow-matching U-Net estimator (decoder params)
    dec_channels: tuple = (256, 256)
    dec_head_dim: int = 64
    dec_heads: int = 2
    dec_n_blocks: int = 1
    dec_mid_blocks: int = 2
    dec_act: str = "snakebeta"
    sigma_min: float = 1e-4
    solver: str = "euler"
    mel_mean: float = -6.630575
    mel_std: float = 2.482914

    @property
    def enc_hidden(self) -> int:  # text_encoder.py:365
        return self.enc_channels + (self.spk_emb_dim if self.n_spks > 1 else 0)

    @property
    def dec_in(self) -> int:  # flow_matching.py:130  (x | mu | spk)
        return 2 * self.n_feats + (self.spk_emb_dim if self.n_spks > 1 else 0)

    @property
    def time_dim(self) -> int:  # decoder.py:221
        return self.dec_channels[0] * 4

    def constructor_kwargs(self) -> dict:
        """kwargs in the shape MatchaTTS.__init__ takes them (matcha_tts.py:27-42)."""
        enc = AttrDict(
            encoder_type="RoPE Encoder",
            encoder_params=AttrDict(
                n_feats=self.n_feats, n_channels=self.enc_channels, filter_channels=self.enc_filter_channels,
                filter_channels_dp=self.enc_filter_channels_dp, n_heads=self.enc_heads, n_layers=self.enc_layers,
                kernel_size=self.enc_kernel, p_dropout=0.1, spk_emb_dim=self.spk_emb_dim, n_spks=1,
                prenet=self.enc_prenet),
            duration_predictor_params=AttrDict(
                filter_channels_dp=self.enc_filter_channels_dp, kernel_size=3, p_dropout=0.1),
        )
        dec = AttrDict(channels=list(self.dec_channels), dropout=0.05, attention_head_dim=self.dec_head_dim,
                       n_blocks=self.dec_n_blocks, num_mid_blocks=self.dec_mid_blocks, num_heads=self.dec_heads,
                       act_fn=self.dec_act)
        cfm = AttrDict(name="CFM", solver=self.solver, sigma_min=self.sigma_min)
        return dict(n_vocab=self.n_vocab, n_spks=self.n_spks, spk_emb_dim=self.spk_emb_dim, n_feats=self.n_feats,
                    encoder=enc, decoder=dec, cfm=cfm,
                    data_statistics=AttrDict(mel_mean=self.mel_mean, mel_std=self.mel_std), out_size=None)

    @staticmethod
    def from_constructor_kwargs(n_vocab, n_spks, spk_emb_dim, n_feats, encoder, decoder, cfm, data_statistics,
                                **_unused) -> "MatchaConfig":
        ep = encoder["encoder_params"] if isinstance(encoder, dict) else encoder.encoder_params
        g = (lambda o, k, d=None: o.get(k, d) if isinstance(o, dict) else getattr(o, k, d))
        ds = data_statistics or {"mel_mean": 0.0, "mel_std": 1.0}
        return MatchaConfig(
            n_vocab=n_vocab, n_spks=n_spks, spk_emb_dim=spk_emb_dim, n_feats=n_feats,
            enc_channels=g(ep, "n_channels"), enc_filter_channels=g(ep, "filter_channels"),
            enc_filter_channels_dp=g(ep, "filter_channels_dp"), enc_heads=g(ep, "n_heads"),
            enc_layers=g(ep, "n_layers"), enc_kernel=g(ep, "kernel_size"), enc_prenet=bool(g(ep, "prenet", True)),
            dec_channels=tuple(g(decoder, "channels")), dec_head_dim=g(decoder, "attention_head_dim"),
            dec_heads=g(decoder, "num_heads"), dec_n_blocks=g(decoder, "n_blocks"),
            dec_mid_blocks=g(decoder, "num_mid_blocks"), dec_act=g(decoder, "act_fn"),
            sigma_min=float(g(cfm, "sigma_min", 1e-4)), solver=g(cfm, "solver", "euler"),
            mel_mean=float(g(ds, "mel_mean")), mel_std=float(g(ds, "mel_std")))


VCTK = MatchaConfig()

# matcha/hifigan/config.py:1-28 (only the generator-relevant keys)
HIFIGAN_V1 = AttrDict(
    resblock="1",
    upsample_rates=[8, 8, 2, 2],
    upsample_kernel_sizes=[16, 16, 4, 4],
    upsample_initial_channel=512,
    resblock_kernel_sizes=[3, 7, 11],
    resblock_dilation_sizes=[[1, 3, 5], [1, 3, 5], [1, 3, 5]],
    num_mels=80,
    n_fft=1024,
    hop_size=256,
    win_size=1024,
    sampling_rate=22050,
)

SAMPLE_RATE = 22050
HOP = 256

# feel_me.py:84-96 (female voice map) and feel_me.py:99-111 / case_studies/case3_game/main.py:111-123 (male map)
EMOJI_MAPPING_FEMALE = {
    "\U0001F60D": 107, "\U0001F621": 58, "\U0001F60E": 79, "\U0001F62D": 103, "\U0001F644": 66, "\U0001F601": 18,
    "\U0001F642": 12, "\U0001F923": 15, "\U0001F62E": 54, "\U0001F605": 22, "\U0001F914": 17,
}
EMOJI_MAPPING_MALE = {
    "\U0001F60D": 4, "\U0001F621": 5, "\U0001F60E": 6, "\U0001F62D": 13, "\U0001F644": 16, "\U0001F601": 26,
    "\U0001F642": 30, "\U0001F923": 38, "\U0001F62E": 60, "\U0001F605": 82, "\U0001F914": 97,
}
