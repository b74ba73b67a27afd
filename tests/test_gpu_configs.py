"""GPU parity at the sizes BASELINE.json's configs name (run with -m gpu on the B200 box).  Every test compares the CUDA
path, called through the C ABI, with the CPU oracle on the SAME inputs and prior noise:

  configs[1]  the whole 32-utterance batch, n_timesteps = 10, bf16: mel AND waveform of all 32 items vs the oracle
  configs[2]  real micro-batches of the 1024-utterance mixed-length list (shortest / median / longest), ragged vocoder
  configs[3]  n_timesteps = 50 and length_scale 0.8 / 1.0 / 1.2: the mel, not only the durations
  configs[4]  one 60-second segment (T = 5168 frames) through the vocoder, fp32 and bf16

Tolerances are BASELINE.json's: durations / alignment bit-exact, mel / waveform rel-L2 <= 1e-4 (fp32), <= 1e-2 (bf16)."""
import pytest
import torch

import emojivoice_b200 as ev
from emojivoice_b200 import sharding, synthetic
from emojivoice_b200.batch import collate
from emojivoice_b200.config import HIFIGAN_V1, VCTK
from oracle import hifigan_oracle as ho
from oracle import matcha_oracle as mo
from oracle.metrics import mel_cepstral_distortion
from tests.conftest import rel_l2

pytestmark = pytest.mark.gpu

TOL = {"fp32": 1e-4, "bf16": 1e-2}


@pytest.fixture(scope="module")
def matcha(matcha_sd):
    m = ev.MatchaTTS(**VCTK.constructor_kwargs())
    m.load_state_dict(matcha_sd)
    return m


@pytest.fixture(scope="module")
def hifigan():
    sd = synthetic.hifigan_state_dict(HIFIGAN_V1, seed=4321, gain=1.0)
    g = ev.Generator(HIFIGAN_V1)
    g.load_state_dict(sd)
    g.remove_weight_norm()
    return g, sd


def _oracle_pair(matcha_sd, x, xl, spk, n, ls, seed):
    """Oracle synthesis with seeded prior noise -> (ref dict, z)."""
    probe = mo.synthesise(matcha_sd, VCTK, x, xl, 1, 0.667, spk, ls)
    z = synthetic.prior_noise(x.shape[0], 80, probe["t_pad"], seed=seed)
    return mo.synthesise(matcha_sd, VCTK, x, xl, n, 0.667, spk, ls, z=z), z


def _assert_alignment_exact(out, ref):
    assert torch.equal(out["w_ceil"].cpu(), ref["w_ceil"]), "durations differ from the oracle"
    assert out["mel_lengths"].cpu().tolist() == ref["mel_lengths"].tolist()
    assert out["t_pad"] == ref["t_pad"]
    assert torch.equal(out["attn"].cpu(), ref["attn"])


def test_config2_full_batch_bf16_mel_and_waveform_match_the_oracle(matcha, matcha_sd, hifigan, capsys):
    """All 32 utterances of the bench batch (seed 2000, n_timesteps 10, T 0.667, length_scale 0.8): the tensor-core path's mel
    and waveform against the fp32 oracle run on the same batch and noise (about 20 s of CPU work), dense and ragged vocoder."""
    gen, hsd = hifigan
    x, xl, spk = synthetic.phoneme_batch(32, 60, 90, seed=2000)
    ref, z = _oracle_pair(matcha_sd, x, xl, spk, 10, 0.8, seed=7)
    ref_wav = ho.generator(hsd, HIFIGAN_V1, ref["mel"])
    out = matcha.synthesise(x, xl, 10, 0.667, spk, 0.8, z=z, dtype="bf16")
    _assert_alignment_exact(out, ref)
    e_mel = rel_l2(out["mel"].cpu(), ref["mel"])
    e_full = rel_l2(out["decoder_outputs_full"].cpu(), ref["decoder_outputs_full"])
    mcd = mel_cepstral_distortion(out["mel"], ref["mel"], ref["mel_lengths"])
    wav = gen(out["mel"], dtype="bf16")
    e_wav = rel_l2(wav.cpu(), ref_wav)
    lens = ref["mel_lengths"].tolist()
    rag = gen(out["mel"], dtype="bf16", lengths=out["mel_lengths"])
    e_rag = max(rel_l2(rag[b, :, : n * 256].cpu(), ref_wav[b, :, : n * 256]) for b, n in enumerate(lens))
    with capsys.disabled():
        print(f"\nconfig 2 (32 x ~5 s, n=10, bf16) vs oracle: mel rel-L2 {e_mel:.2e} (padded extent {e_full:.2e}), MCD {mcd:.2e} dB, "
              f"waveform rel-L2 {e_wav:.2e}, worst cropped ragged item {e_rag:.2e}")
    assert e_mel < TOL["bf16"] and e_full < TOL["bf16"]
    assert e_wav < TOL["bf16"] and e_rag < TOL["bf16"]
    # per item too: no utterance hides behind the batch norm
    for b, n in enumerate(lens):
        assert rel_l2(out["mel"][b, :, :n].cpu(), ref["mel"][b, :, :n]) < TOL["bf16"], b
    # fp32 (parity) mode on the same batch
    out32 = matcha.synthesise(x, xl, 10, 0.667, spk, 0.8, z=z, dtype="fp32")
    assert rel_l2(out32["mel"].cpu(), ref["mel"]) < TOL["fp32"]
    sel = [0, 31]
    assert rel_l2(gen(out32["mel"][sel], dtype="fp32").cpu(), ho.generator(hsd, HIFIGAN_V1, out32["mel"][sel].cpu())) < TOL["fp32"]


@pytest.mark.parametrize("ls", [0.8, 1.0, 1.2])
def test_config4_fifty_euler_steps_and_length_scales_match_the_oracle(matcha, matcha_sd, ls):
    """ODE sweep: n_timesteps = 50 (and 2) at three length scales, B = 4 -- mel against the oracle, durations bit-exact."""
    x, xl, spk = synthetic.phoneme_batch(4, 15, 40, seed=400 + int(ls * 10))
    for n in (50, 2):
        ref, z = _oracle_pair(matcha_sd, x, xl, spk, n, ls, seed=50 + n)
        for prec in ("fp32", "bf16"):
            out = matcha.synthesise(x, xl, n, 0.667, spk, ls, z=z, dtype=prec)
            _assert_alignment_exact(out, ref)
            assert rel_l2(out["mel"].cpu(), ref["mel"]) < TOL[prec], (n, ls, prec)
            assert rel_l2(out["decoder_outputs_full"].cpu(), ref["decoder_outputs_full"]) < TOL[prec], (n, ls, prec)


@pytest.mark.parametrize("prec", ["fp32", "bf16"])
def test_config5_sixty_second_segment_matches_the_oracle(hifigan, prec):
    """Vocoder only, one 60-second segment (T = 5168 frames, 1.32 M samples): thousands of tiles per item in the last stages,
    dense and ragged (a second, shorter item), against the oracle's generator."""
    gen, hsd = hifigan
    T = int(round(60 * 22050 / 256))
    assert T == 5168
    mel = synthetic.synthetic_mel(1, T, seed=60)
    ref = ho.generator(hsd, HIFIGAN_V1, mel)
    wav = gen(mel, dtype=prec)
    assert wav.shape == ref.shape == (1, 1, T * 256)
    assert rel_l2(wav.cpu(), ref) < TOL[prec]
    # ragged: the same segment next to a 10-second neighbour
    mel2 = torch.cat([mel, mel.flip(2)], 0)
    rag = gen(mel2, dtype=prec, lengths=[T, 861])
    assert torch.equal(rag[0], wav[0])
    ref_b = ho.generator(hsd, HIFIGAN_V1, mel2[1:2, :, : 861 + 16])[:, :, : 861 * 256]   # 13-frame receptive field + slack
    assert rel_l2(rag[1:2, :, : 861 * 256].cpu(), ref_b) < TOL[prec]
    assert float(rag[1, :, 861 * 256:].abs().sum()) == 0.0


def test_config3_real_microbatches_match_the_oracle(matcha_sd, hifigan, capsys):
    """The 1024-utterance mixed-length list of config 3, cut by `sharding.microbatches` exactly as `synthesise_corpus` does:
    the shortest, the median and the longest micro-batch (32 utterances each) go through the corpus driver (bf16, ragged
    vocoder, copy-stream read-back, crop) and through the oracle with the same prior noise; every utterance's cropped
    waveform must match (cli.py:307-311)."""
    gen, hsd = hifigan
    model = ev.MatchaTTS(**VCTK.constructor_kwargs(), precision="bf16")
    model.load_state_dict(matcha_sd)
    utts = synthetic.mixed_length_corpus(1024)
    plan = sharding.microbatches([len(u[0]) for u in utts], 32, n_timesteps=10)
    assert len(plan) == 32
    worst = {}
    for pick in (0, len(plan) // 2, len(plan) - 1):
        sub = [utts[i] for i in plan[pick].items]
        ref_wavs, ref_len = {}, {}

        def z_fn(mb, model_, x, xl, spks):
            ref, z = _oracle_pair(matcha_sd, x, xl, spks, 10, 0.8, seed=300 + pick)
            wav = ho.generator(hsd, HIFIGAN_V1, ref["mel"]).clamp(-1, 1)
            for j, i in enumerate(mb.items):
                n = int(ref["mel_lengths"][j])
                ref_wavs[i], ref_len[i] = wav[j, 0, : n * 256], n
            return z

        # the sub-list is already sorted by length, so the driver forms the same single micro-batch
        res, stats = ev.synthesise_corpus(model, gen, sub, batch_size=32, n_timesteps=10, temperature=0.667, length_scale=0.8, z_fn=z_fn)
        assert sorted(res) == list(range(32)) and stats.utterances == 32
        errs = []
        for i in range(32):
            assert res[i]["mel_length"] == ref_len[i]
            assert res[i]["waveform"].shape == ref_wavs[i].shape
            errs.append(rel_l2(res[i]["waveform"], ref_wavs[i]))
        worst[pick] = max(errs)
        assert max(errs) < TOL["bf16"], (pick, max(errs))
    with capsys.disabled():
        print("\nconfig 3 micro-batches (shortest / median / longest), worst cropped-waveform rel-L2 vs oracle:",
              ", ".join(f"#{k}: {v:.2e}" for k, v in worst.items()))


def test_narrow_decoder_with_poisoned_workspace(capsys):
    """dec_channels = 128 takes the unfused transformer path (separate LayerNorm / FF convs, multiplicative-mask epilogues in
    round 1).  A ragged bf16 batch on a workspace pre-filled with 0xFF (NaN as bf16 / fp32) must still match the oracle:
    rows a skipped producer leaves unwritten may never surface as NaN * 0."""
    import dataclasses

    cfg = dataclasses.replace(VCTK, dec_channels=(128, 128))
    sd = synthetic.matcha_state_dict(cfg, seed=77)
    model = ev.MatchaTTS(**cfg.constructor_kwargs(), precision="bf16", cuda_graphs=False)
    model.load_state_dict(sd)
    x, xl, spk = synthetic.phoneme_batch(5, 3, 70, seed=78)
    xl[1] = 3                                                   # one very short item: whole tiles of padding
    x[1, 3:] = 0
    probe = mo.synthesise(sd, cfg, x, xl, 1, 0.667, spk, 1.0)
    z = synthetic.prior_noise(5, 80, probe["t_pad"], seed=79)
    ref = mo.synthesise(sd, cfg, x, xl, 3, 0.667, spk, 1.0, z=z)
    model.synthesise(x, xl, 3, 0.667, spk, 1.0, z=z)            # sizes the workspace
    with torch.inference_mode():
        model._ctx._ws.fill_(0xFF)
    out = model.synthesise(x, xl, 3, 0.667, spk, 1.0, z=z)
    assert torch.isfinite(out["mel"]).all()
    _assert_alignment_exact(out, ref)
    err = rel_l2(out["mel"].cpu(), ref["mel"])
    with capsys.disabled():
        print(f"\ndec_channels=128 (unfused path), poisoned workspace: mel rel-L2 {err:.2e}")
    assert err < TOL["bf16"]
    assert rel_l2(out["decoder_outputs_full"].cpu(), ref["decoder_outputs_full"]) < TOL["bf16"]


def test_synthesise_file_script_to_wav_on_disk(matcha, matcha_sd, hifigan, tmp_path):
    """Row f2 composed (cli.py:226-250, :277-317, :129-135): a script file with the three line formats goes through cleaner rules ->
    (stand-in) phonemiser -> symbol ids -> blank interspersing -> sharded micro-batches -> synthesise -> ragged vocoder -> crop ->
    PCM_24 WAV + NPY on disk in one call; the files must hold exactly what the oracle synthesises for the same ids (the model
    runs fp32 here and draws its own prior noise, so the comparison uses temperature 0: z plays no role)."""
    from emojivoice_b200 import audio_io, text_cleaners, text_frontend

    gen, hsd = hifigan
    fake_g2p = {"doctor who is here.": "dˈɑktɚ hˈuː ɪz hˈɪɹ.", "i am so happy today !": "aɪ æm sˈoʊ hˈæpi tədˈeɪ !", "plain line": "plˈeɪn lˈaɪn"}
    script = tmp_path / "script.txt"
    script.write_text("Dr. Who is here.|107\n\nI am so happy today \U0001F601!\nPlain   line\n", encoding="utf-8")
    model = ev.MatchaTTS(**VCTK.constructor_kwargs(), precision="fp32")
    model.load_state_dict(matcha_sd)
    written, stats = ev.synthesise_file(model, gen, script, tmp_path / "out", phonemizer=lambda t: fake_g2p[" ".join(t.split())], n_timesteps=2,
                                        temperature=0.0, length_scale=1.0, batch_size=1, spk=12)   # batch of 1: the decoder is not padding-invariant (H1)
    assert [w[0] for w in written] == [0, 1, 2] and stats.utterances == 3
    happy = ev.EMOJI_MAPPING_FEMALE["\U0001F601"]
    names = ["utterance_000_speaker_107", f"utterance_001_speaker_{happy:03d}", "utterance_002_speaker_012"]
    for (i, path, n), name, spk, key in zip(written, names, (107, happy, 12), fake_g2p):
        assert path.endswith(name + ".wav")
        ids = text_frontend.process_text("x", lambda _t, _k=key: text_cleaners.collapse_whitespace(fake_g2p[_k]))["x"]
        ref = mo.synthesise(matcha_sd, VCTK, ids, torch.tensor([ids.shape[1]]), 2, 0.0, torch.tensor([spk]), 1.0)
        assert n == int(ref["mel_lengths"][0])
        wav_ref = ho.generator(hsd, HIFIGAN_V1, ref["mel"][:, :, :n + 16])[0, 0, : n * 256].clamp(-1, 1)
        wav, sr = audio_io.read_wav_pcm24(path)
        assert sr == 22050 and wav.shape == (n * 256,)
        assert rel_l2(torch.from_numpy(wav), wav_ref) < TOL["bf16"]              # the vocoder fixture runs its default bf16 mode
        mel = torch.from_numpy(__import__("numpy").load(path[:-4] + ".npy"))
        assert mel.shape == (80, n) and rel_l2(mel, ref["mel"][0, :, :n]) < TOL["fp32"]
