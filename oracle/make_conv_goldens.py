"""Generates tests/golden/convblock.npz from the UNMODIFIED reference ConvolutionBlock (run in the build container only):
    python -m oracle.make_conv_goldens
TEST INFRASTRUCTURE ONLY."""
import os

import numpy as np
import torch

from oracle import conv_oracle as CO
from oracle import ref_import

SEED, D, K, T = 4321, 512, 31, 48


def main():
    ref_import.load_reference()
    from lightspeech.layers.block import ConvolutionBlock              # block.py:129-171
    W = CO.make_conv_weights(SEED, D, K)
    blk = ConvolutionBlock(D, K, 0.1).eval()
    missing = blk.load_state_dict({k: torch.from_numpy(v) for k, v in W.items()}, strict=False)
    assert not missing.missing_keys or missing.missing_keys == ["norm.num_batches_tracked"], missing
    rng = np.random.Generator(np.random.PCG64(SEED + 1))
    x = rng.standard_normal((T, D)).astype(np.float32)
    with torch.no_grad():
        y = blk(torch.from_numpy(x)[None], torch.zeros(1, T, dtype=torch.bool))[0].numpy()
    out = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests", "golden", "convblock.npz")
    np.savez_compressed(out, x=x, y=y, seed=SEED, d=D, k=K)
    print(out, y.shape, float(np.abs(y - CO.conv_block_full(x, W)).max()))


if __name__ == "__main__":
    main()
