"""Seeded inputs of the hot-kernel fixture (``hot_kernels_reference.npz``): regenerated identically
by the generator (run against the unmodified reference) and by the tests."""
import numpy as np


def hot_inputs():
    rng = np.random.RandomState(4242)

    def randn(*shape):
        return rng.randn(*shape).astype(np.float32)

    inp = {}
    inp['X2'] = randn(8192, 64) * 1.2 + 0.3                      # cfg2
    inp['X3'] = randn(4096, 16) + randn(8, 16)[rng.randint(8, size=4096)] * 2.0     # cfg3 data
    logits = randn(4096, 8) * 2.0
    r = np.exp(logits - logits.max(1, keepdims=True))
    inp['R3'] = (r / r.sum(1, keepdims=True)).astype(np.float32)
    inp['Lg3'] = (randn(512, 128) * 3.0).astype(np.float32)
    inp['X4'] = randn(2048, 256)                                  # cfg4
    inp['t4'] = (inp['X4'] @ (randn(256) / 16.0) + 0.1 * randn(2048)).astype(np.float32)
    inp['X5'] = randn(2048, 128)                                  # cfg5
    inp['W5'] = (randn(64, 128) / np.sqrt(128.0)).astype(np.float32)
    inp['b5'] = (rng.rand(2048) < 0.5).astype(np.float32)
    # cfg3 at K = 256 (the extent the pre-split / CTA-pair statistics route serves); appended last so that the
    # inputs above keep their values
    inp['X3b'] = randn(2048, 16) + randn(256, 16)[rng.randint(256, size=2048)] * 2.0
    inp['Lg3b'] = (randn(2048, 256) * 2.5).astype(np.float32)
    inp['X4b'] = randn(1536, 136) * 0.8 + 0.1                     # cfg4 at a D the Gram kernel serves through zero padding
    return {k: np.ascontiguousarray(v, dtype=np.float32) for k, v in inp.items()}
