"""CPU model of the error-compensated BF16 arithmetic of the tensor-pipe kernels and of the
two-product diagonal blocks of ``gram_pair_kernel`` (csrc/gram_sm100.cu; it evaluates what the reference
computes as ``_tensordot(_dimshuffle(X,1,0), X, [1],[0])``, bayesic/algebra.py:527-551 -> 1347-1351).

numpy emulation of the operand split (round-to-nearest-even to bfloat16, exact products, float64 sums), so the
identities the kernel relies on are pinned without a GPU:
  x = b1 + b2 + O(2^-17 x);  2 b2 is a bfloat16;  S = b1^T b1 + b1^T (2 b2)  =>  (S + S^T) / 2 = b1^T b1 + b1^T b2 + b2^T b1.
"""
import numpy as np
import pytest


def bf16(x):
    """float32 -> nearest bfloat16 (ties to even), returned as float32"""
    x = np.ascontiguousarray(x, dtype=np.float32)
    u = x.view(np.uint32).astype(np.uint64)
    r = ((u + 0x7FFF + ((u >> 16) & 1)) >> 16) << 16
    return r.astype(np.uint32).view(np.float32)


def is_bf16(x):
    return np.array_equal(bf16(x), np.asarray(x, np.float32))


def split(x):
    b1 = bf16(x)
    b2 = bf16(np.asarray(x, np.float32) - b1)
    return b1, b2


@pytest.fixture(scope='module')
def data():
    rng = np.random.RandomState(1234)
    return (rng.randn(4096, 96) * np.exp(rng.randn(96))[None, :] + rng.randn(96)[None, :]).astype(np.float32)


def test_two_part_split_residual(data):
    b1, b2 = split(data)
    resid = data.astype(np.float64) - b1.astype(np.float64) - b2.astype(np.float64)
    # b1 keeps 8 significant bits, b2 the next 8: what is left is below 2^-16 |x| (2^-17 typical)
    assert np.all(np.abs(resid) <= 2.0 ** -16 * np.abs(data).astype(np.float64))
    assert np.sqrt(np.mean((resid / np.maximum(np.abs(data), 1e-30)) ** 2)) < 2.0 ** -17


def test_twice_the_low_part_is_a_bfloat16(data):
    _, b2 = split(data)
    assert is_bf16(2.0 * b2)                       # a power-of-two scale only moves the exponent
    assert np.array_equal((2.0 * b2).astype(np.float64), 2.0 * b2.astype(np.float64))


def test_diagonal_block_two_products_equal_three(data):
    b1, b2 = split(data)
    b1, b2 = b1.astype(np.float64), b2.astype(np.float64)
    three = b1.T @ b1 + b1.T @ b2 + b2.T @ b1            # an off-diagonal block's products, on the diagonal
    s = b1.T @ b1 + b1.T @ (2.0 * b2)                    # what the issuer accumulates: P + 2 Q
    two = 0.5 * (s + s.T)                                # the finalize kernel
    np.testing.assert_allclose(two, three, rtol=1e-13, atol=1e-13 * np.abs(three).max())
    assert np.array_equal(two, two.T)                    # exactly symmetric, as the finalize writes it


def test_three_products_against_float64_and_float32(data):
    """The compensated Gram is as accurate as a float32 BLAS on well-scaled data: the dropped b2^T b2 term and the
    split residual are ~2^-16 relative per product (DESIGN.md 4.3 explains where this stops holding: entries that
    cancel to far below the column norms)."""
    x64 = data.astype(np.float64)
    ref = x64.T @ x64
    b1, b2 = split(data)
    b1, b2 = b1.astype(np.float64), b2.astype(np.float64)
    got = b1.T @ b1 + b1.T @ b2 + b2.T @ b1
    scale = np.sqrt(np.outer(np.diag(ref), np.diag(ref)))        # |x_d| |x_e|: the natural size of an entry
    assert np.max(np.abs(got - ref) / scale) < 2.0 ** -15
    f32 = (data.T @ data).astype(np.float64)
    assert np.max(np.abs(f32 - ref) / scale) < 2.0 ** -15          # the reference's own arithmetic, same bar
