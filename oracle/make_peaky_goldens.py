"""A checkpoint with PEAKED posteriors + its goldens from the unmodified reference.  TEST INFRASTRUCTURE ONLY.
Run in the build container:  python -m oracle.make_peaky_goldens

Why: the seeded random-init model has almost flat posteriors (top-1 / top-2 log-prob gap: min 2e-4, median 0.04), so a
bf16 path within the north-star tolerance (max-abs 1e-2 on log-probs) cannot also be token-exact on it — neither can the reference's
own bf16 autocast.  Trained CTC models are not like that: their top-1 posterior sits near 1.  Scaling decoder.linear2
does not create such a model (argmax and the margin / error ratio are scale-invariant), so this script FITS the CTC output layer:
on the fixture audio it takes the hidden activations h_t = SiLU(linear1(encoder output)) of the reference model (weights seed 1234)
and the model's own greedy path c_t, and solves the ridge regression  W (h_t - mean h) + b = TARGET * onehot(c_t)  in float64 (SVD of the
centred activations; the random-init model's frames are nearly collinear, so the ridge term keeps the weights — and with them the
amplification of upstream rounding errors — as small as the wanted margins allow).  Everything else (encoder, linear1) keeps the seeded
weights.  The result decodes the SAME token sequences as the random-init model with top-2 log-prob margins >= 10 and a top-1 posterior
> 0.99 — the regime a deployed checkpoint is in — and the GPU test asserts that the bf16 (FAST) engine is 100 % token-exact on it.  The fitted layer is stored (tests/golden/peaky_head.npz) because a re-fit on another box could differ in the
last bit; the goldens (tests/golden/peaky_*.npz) are the unmodified reference's outputs with that checkpoint."""
from __future__ import annotations

import json
import os

import numpy as np
import torch

from oracle import lightspeech_oracle as O
from oracle import ref_import
from oracle.make_goldens import GOLD, WEIGHT_SEED, run_reference

TARGET, RIDGE = 25.0, 1e-2
CASES = ["synth_noise", "synth_tone", "testwav"]


def main():
    torch.set_num_threads(max(1, os.cpu_count() or 1))
    geo = O.CANONICAL
    W = O.make_weights(WEIGHT_SEED, geo)
    R, m = ref_import.build_reference_model(W, geo)
    hidden = []
    hook = m.decoder.linear2.register_forward_hook(lambda mod, inp, out: hidden.append(inp[0].detach()[0].double().numpy().copy()))
    targets, pcms = [], {}
    for name in CASES:
        pcm = np.load(os.path.join(GOLD, f"{name}.npz"))["pcm"]
        pcms[name] = pcm
        out = run_reference(R, m, pcm, geo)
        targets.append(out["argmax"].reshape(-1))
    hook.remove()
    H = np.concatenate(hidden)                                             # [T, 512] float64
    c = np.concatenate(targets)
    T = H.shape[0]
    assert c.shape[0] == T
    Y = np.zeros((T, geo.vocab))
    Y[np.arange(T), c] = TARGET
    mu, ybar = H.mean(0), Y.mean(0)
    U, S, Vt = np.linalg.svd(H - mu, full_matrices=False)
    Wf = (Vt.T * (S / (S * S + RIDGE))) @ (U.T @ (Y - ybar))                # [512, V]
    w2, b2 = Wf.T.astype(np.float32), (ybar - mu @ Wf).astype(np.float32)
    z = H @ w2.astype(np.float64).T + b2
    s = np.sort(z, axis=1)
    print(f"fitted on {T} frames: min top-2 logit gap {np.min(s[:, -1] - s[:, -2]):.3f}, |W| row norm max {np.linalg.norm(w2, axis=1).max():.2f} "
          f"(seeded layer: {np.linalg.norm(W['decoder.linear2.weight'], axis=1).max():.2f}), argmax kept: {(z.argmax(1) == c).all()}")
    assert (z.argmax(1) == c).all() and np.min(s[:, -1] - s[:, -2]) > 10.0
    np.savez_compressed(os.path.join(GOLD, "peaky_head.npz"), weight=w2, bias=b2)
    Wp = dict(W)
    Wp["decoder.linear2.weight"], Wp["decoder.linear2.bias"] = w2, b2
    R, mp = ref_import.build_reference_model(Wp, geo)
    meta_p = os.path.join(GOLD, "meta.json")
    with open(meta_p, encoding="utf-8") as f:
        meta = json.load(f)
    meta["peaky"] = {"target": TARGET, "ridge": RIDGE, "fit_cases": CASES, "cases": {}}
    for name in CASES:
        out = run_reference(R, mp, pcms[name], geo)
        e = out["emission"]
        s = np.sort(e, axis=-1)
        np.savez_compressed(os.path.join(GOLD, f"peaky_{name}.npz"), emission=e.astype(np.float32), argmax=out["argmax"], last_blank=out["last_blank"])
        meta["peaky"]["cases"][name] = {"texts": out["texts"], "n_chunks": int(e.shape[0]), "min_margin": float((s[..., -1] - s[..., -2]).min()),
                                        "top1_prob_min": float(np.exp(s[..., -1]).min())}
        print(name, e.shape, "min margin", meta["peaky"]["cases"][name]["min_margin"], "min top-1 prob", meta["peaky"]["cases"][name]["top1_prob_min"])
    with open(meta_p, "w", encoding="utf-8") as f:
        json.dump(meta, f, ensure_ascii=False, indent=0)


if __name__ == "__main__":
    main()
