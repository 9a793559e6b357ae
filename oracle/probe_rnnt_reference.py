"""TEST INFRASTRUCTURE / evidence only.  Does the reference's English path (EmformerRNNT.stream, recognition.py:96-133) run with the
chunk geometry its own configuration gives it (config/asr-online-en.yaml:68-74, ``audio_en``: segment_size 8, context_size 4, bias 0)?

    python oracle/probe_rnnt_reference.py

EmformerRNNT.stream = feature extractor (MelSpectrogram(16 kHz, n_fft 400, 80 mels, hop 160) -> piecewise-linear log -> global stats)
-> RNNTBeamSearch(emformer_rnnt_base(4097), blank 4096).infer(features, length, beam_width=10, state, hypothesis).  The checkpoint, the
global-stats JSON and the SentencePiece model are not in the repo, so random-init weights stand in (the check below is about shapes).
With torchaudio 2.11 (the reference pins no version, Dockerfile:39):
  audio_en as committed (segment_size 8)  -> chunk 1920 samples -> 13 feature frames -> 3 rows after the x4 time reduction
                                             -> Emformer.infer raises ValueError (it expects segment 4 + right context 1 = 5 rows)
  segment_size 16 ("Reduced from 16")     -> chunk 3200 samples -> 21 frames -> 5 rows -> runs
i.e. the English path cannot process a single chunk as configured; DESIGN.md section 7 records this as the reason it is out of scope."""
import torch
import torchaudio
from torchaudio.models import RNNTBeamSearch, emformer_rnnt_base
from torchaudio.pipelines.rnnt_pipeline import _gain, _piecewise_linear_log


def main() -> None:
    torch.manual_seed(0)
    dec = RNNTBeamSearch(emformer_rnnt_base(num_symbols=4097).eval(), blank=4096)
    mel = torchaudio.transforms.MelSpectrogram(sample_rate=16000, n_fft=400, n_mels=80, hop_length=160)
    print("torchaudio", torchaudio.__version__)
    for seg, ctx, bias in ((8, 4, 0), (16, 4, 0)):
        n = (seg + ctx + bias) * 160                       # utils.py:17-22 (AudioConfig.chunk_length)
        feats = _piecewise_linear_log(mel(0.1 * torch.randn(n)).transpose(1, 0) * _gain)
        try:
            with torch.inference_mode():
                hypos, _ = dec.infer(feats, torch.tensor([feats.shape[0]]), 10, state=None, hypothesis=None)
            print(f"segment_size {seg}: chunk {n} samples -> features {tuple(feats.shape)} -> ok, {len(hypos)} hypotheses")
        except Exception as ex:                           # noqa: BLE001
            print(f"segment_size {seg}: chunk {n} samples -> features {tuple(feats.shape)} -> {type(ex).__name__}: {ex}")


if __name__ == "__main__":
    main()
