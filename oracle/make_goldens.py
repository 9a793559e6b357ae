"""Generate tests/golden/* from the UNMODIFIED reference imported in the build container.
TEST INFRASTRUCTURE ONLY.  Run:  python -m oracle.make_goldens   (needs /root/reference)

The reference ships no golden vectors for this path, so these fixtures are
"outputs of the reference itself run here": ``LightningASR.stream`` / ``greedy_search``
(recognition.py:191-204, :33-57) at batch 1 per stream, ``extract_filterbank`` (audio.py:9-30),
and ``torchaudio.compliance.kaldi.fbank`` for the north-star Kaldi front-end.
Weights are regenerated from a seed on every box (oracle.make_weights), not stored.
"""
from __future__ import annotations

import json
import os
import sys
import wave

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from oracle import lightspeech_oracle as O  # noqa: E402
from oracle import ref_import  # noqa: E402

GOLD = os.path.join(ROOT, "tests", "golden")
WEIGHT_SEED = 1234


def synth_audio(seed: int, n: int, sigma: float = 0.1, tone: float = 0.0) -> np.ndarray:
    """int16 PCM: round(N(0, (sigma*32768)^2)) + optional 440 Hz tone, clipped."""
    rng = np.random.Generator(np.random.PCG64(seed))
    x = sigma * rng.standard_normal(n)
    if tone:
        x = x + tone * np.sin(2 * np.pi * 440.0 * np.arange(n) / 16000.0)
    return np.clip(np.round(x * 32768.0), -32768, 32767).astype(np.int16)


def pcm_to_float(p: np.ndarray) -> np.ndarray:
    """streaming_server.py:362-363."""
    return (p.astype(np.float32) / np.float32(32768.0)).astype(np.float32)


def load_test_wav_16k() -> np.ndarray:
    """test.wav (44.1 kHz mono s16) -> 16 kHz int16.  The reference's own resampler (pydub/audioop)
    is absent; torchaudio.functional.resample stands in and the RESULT is the fixture."""
    import torchaudio
    with wave.open(os.path.join(ref_import.REFERENCE_ROOT, "test.wav"), "rb") as w:
        sr, ch, n = w.getframerate(), w.getnchannels(), w.getnframes()
        raw = np.frombuffer(w.readframes(n), dtype=np.int16).reshape(-1, ch)[:, 0]
    x = torch.from_numpy(raw.astype(np.float32) / 32768.0)
    y = torchaudio.functional.resample(x, sr, 16000)
    return np.clip(np.round(y.numpy() * 32768.0), -32768, 32767).astype(np.int16)


def run_reference(R, m, pcm_i16: np.ndarray, geo, reset_before=(), skip=()):
    """Drive the reference exactly like handle_connection_impl does (streaming_server.py:371-435),
    batch 1.  ``reset_before``: chunk indices before which the state is re-initialised (endpoint,
    :514-515,:530).  ``skip``: chunk indices not sent to the model (VAD gate, :377-379)."""
    audio = torch.from_numpy(pcm_to_float(pcm_i16))
    buf = torch.cat([torch.zeros(geo.buffer_length), audio])          # stream.py:23
    state = m.init_state()
    emission = torch.zeros(0, geo.vocab)
    per_chunk, texts, blanks, argmax = [], [], [], []
    k = 0
    while buf.numel() >= geo.chunk_length:
        if k in reset_before:
            state = m.init_state()
            emission = torch.zeros(0, geo.vocab)
        if k not in skip:
            em, ln, sts = m.stream([buf[None, :geo.chunk_length]], 16000, [state])
            assert int(ln[0]) == geo.seg_rows
            state = sts[0]
            emission = torch.cat([emission, em[0]])                      # :431
            text, last_blank = R.greedy_search(emission)                 # :433
            per_chunk.append(em[0].numpy().copy())
            texts.append(text)
            blanks.append(float(last_blank))
            argmax.append(em[0].argmax(1).numpy().astype(np.int32))
        buf = buf[geo.segment_length:]                                   # stream.py:159
        k += 1
    return dict(emission=np.stack(per_chunk), texts=texts, last_blank=np.asarray(blanks, np.float64),
                argmax=np.stack(argmax),
                state_k_l0=state[0][1][:, 0].numpy().copy(), state_v_l19=state[geo.n_layers - 1][2][:, 0].numpy().copy(),
                past_length=int(state[0][3][0][0]))


def main() -> None:
    os.makedirs(GOLD, exist_ok=True)
    torch.set_num_threads(max(1, os.cpu_count() or 1))
    geo = O.CANONICAL
    W = O.make_weights(WEIGHT_SEED, geo)
    R, m = ref_import.build_reference_model(W, geo)
    meta = {"weight_seed": WEIGHT_SEED, "torch": torch.__version__, "vocab": list(R.vocab), "cases": {}}

    def save_case(name, pcm, geo_name="canonical", **kw):
        g = O.CANONICAL if geo_name == "canonical" else O.LOW_LATENCY
        mm = m if geo_name == "canonical" else m_ll
        out = run_reference(R, mm, pcm, g, **kw)
        np.savez_compressed(os.path.join(GOLD, f"{name}.npz"), pcm=pcm, emission=out["emission"],
                            argmax=out["argmax"], last_blank=out["last_blank"],
                            state_k_l0=out["state_k_l0"], state_v_l19=out["state_v_l19"])
        meta["cases"][name] = {"geometry": geo_name, "texts": out["texts"], "past_length": out["past_length"],
                               "reset_before": sorted(kw.get("reset_before", ())), "skip": sorted(kw.get("skip", ())),
                               "n_chunks": int(out["emission"].shape[0])}
        print(name, out["emission"].shape, "past_length", out["past_length"], "text[-1]=", repr(out["texts"][-1][:50]))

    # (i) seeded Gaussian audio sigma=0.1 (+ tone), 6 chunks
    save_case("synth_noise", synth_audio(7, 16000 * 4))
    save_case("synth_tone", synth_audio(8, 16000 * 3, sigma=0.02, tone=0.3))
    # (ii) test.wav resampled to 16 kHz (config #1)
    save_case("testwav", load_test_wav_16k())
    # (iii) edge cases: silence, full-scale square-ish, DC
    n = geo.segment_length * 3
    save_case("edge_silence", np.zeros(n, np.int16))
    fs = np.where((np.arange(n) // 37) % 2 == 0, 32767, -32768).astype(np.int16)
    save_case("edge_fullscale", fs)
    save_case("edge_dc", np.full(n, 12000, np.int16))
    # (iv) reset mid-stream (endpoint) and VAD-skipped chunks
    save_case("seq_reset_skip", synth_audio(9, 16000 * 5), reset_before=(3,), skip=(1, 5))
    # (v) low-latency geometry (segment 32 frames -> T = 12)
    _, m_ll = ref_import.build_reference_model(W, O.LOW_LATENCY)
    save_case("lowlat_noise", synth_audio(10, 16000 * 2), geo_name="lowlat")

    # fbank-only goldens: extract_filterbank on one chunk (audio.py:9-30)
    from lightspeech.datas.audio import extract_filterbank
    pcm = synth_audio(11, geo.chunk_length, sigma=0.05, tone=0.2)
    fb, lens = extract_filterbank(torch.from_numpy(pcm_to_float(pcm))[None], 16000, "cpu")
    assert int(lens[0]) == geo.frames_per_chunk
    # Kaldi fbank (north-star front-end / config #2): TA:compliance/kaldi.py:514
    import torchaudio
    kpcm = synth_audio(12, 10240 + 240, sigma=0.09, tone=0.1)
    kfb = torchaudio.compliance.kaldi.fbank(torch.from_numpy(kpcm.astype(np.float32))[None], num_mel_bins=80,
                                            dither=0.0, sample_frequency=16000.0)
    kfb_cmvn = torchaudio.compliance.kaldi.fbank(torch.from_numpy(kpcm.astype(np.float32))[None], num_mel_bins=80,
                                                 dither=0.0, sample_frequency=16000.0, subtract_mean=True)
    np.savez_compressed(os.path.join(GOLD, "fbank.npz"), melspec_pcm=pcm, melspec128=fb[0].numpy(),
                        kaldi_pcm=kpcm, kaldi80=kfb.numpy(), kaldi80_cmvn=kfb_cmvn.numpy())
    print("fbank", fb.shape, "kaldi", kfb.shape)

    with open(os.path.join(GOLD, "meta.json"), "w", encoding="utf-8") as f:
        json.dump(meta, f, ensure_ascii=False, indent=0)


if __name__ == "__main__":
    main()
