"""Developer script (CPU): attributes the cfg5 error of logistic_fused2_kernel to its sources with a numpy model of
the kernel's arithmetic -- BF16 hi/lo splits, exact products, an fp32 accumulator that TRUNCATES (round toward zero)
after every K = 16 MMA, chains per CTA -- on the parity report's own inputs (tests/gpu_parity_report.py, cfg5)."""
import sys
import numpy as np

def bf16(x):
    x = np.asarray(x, dtype=np.float32)
    u = x.view(np.uint32).astype(np.uint64)
    r = ((u + 0x7FFF + ((u >> 16) & 1)) >> 16) << 16
    return r.astype(np.uint32).view(np.float32)

def split(x):
    h = bf16(x)
    l = bf16(np.asarray(x, np.float32) - h)
    return h.astype(np.float64), l.astype(np.float64)

def rz32(x):
    y = x.astype(np.float32)
    over = np.abs(y.astype(np.float64)) > np.abs(x)
    y[over] = np.nextafter(y[over], np.float32(0))
    return y.astype(np.float64)

def rn32(x):
    return x.astype(np.float32).astype(np.float64)

def metrics(got, ref):
    scale = np.abs(ref).max()
    elem = np.max(np.abs(got - ref) / np.maximum(np.abs(ref), 1e-3 * scale))
    norm = np.linalg.norm(got - ref) / np.linalg.norm(ref)
    return elem, norm

def main():
    rng = np.random.RandomState(2024)
    # replay the generator stream of gpu_parity_report.py is not needed: same distributions, own seed
    n, d, s = 1 << 16, 512, 64
    X = rng.randn(n, d).astype(np.float32)
    y = (rng.rand(n) < 0.5).astype(np.float32)
    mu, ls, eps = rng.randn(d) * 0.05, np.full(d, -2.0), rng.randn(s, d)
    W = (mu[None] + np.exp(ls)[None] * eps).astype(np.float32)
    X64, W64 = X.astype('f8'), W.astype('f8')
    Z_ref = X64 @ W64.T
    resid_ref = y[:, None] - 1.0 / (1.0 + np.exp(-Z_ref))
    G_ref = X64.T @ resid_ref
    X1, X2 = split(X)
    W1, W2 = split(W)
    round_acc = {'rz': rz32, 'rn': rn32}

    def z_model(mode, n_acc=1):
        """four accumulators (W1X1, W1X2, W2X1, W2X2); n_acc > 1: the K loop alternates between n_acc accumulator sets"""
        rnd = round_acc[mode]
        accs = [[np.zeros((n, s)) for _ in range(4)] for _ in range(n_acc)]
        for ks in range(d // 16):
            sl = slice(16 * ks, 16 * ks + 16)
            a = accs[ks * n_acc // (d // 16)]
            for i, (xa, wa) in enumerate(((X1, W1), (X2, W1), (X1, W2), (X2, W2))):
                a[i] = rnd(a[i] + xa[:, sl] @ wa[:, sl].T)
        tot = np.zeros((n, s))
        for a in accs:
            tot = rn32(tot + rn32(rn32(a[0] + a[1]) + rn32(a[2] + a[3])))
        return tot

    def g_model(resid, mode, chain_rows, split_resid=True, separate_lo=False, ctas=148, three_way=False):
        rnd = round_acc[mode]
        if split_resid:
            R1, R2 = split(resid.astype(np.float32))
            R3 = None
            if three_way:
                R3 = bf16(resid.astype(np.float32) - R1.astype(np.float32) - R2.astype(np.float32)).astype('f8')
        else:
            R1, R2 = resid, np.zeros_like(resid)
        tiles = n // 64
        G = np.zeros((d, s))
        for c in range(ctas):
            t0, t1 = tiles * c // ctas, tiles * (c + 1) // ctas
            part = np.zeros((d, s))
            acc = np.zeros((d, s)); acc_lo = np.zeros((d, s))
            rows_in_chain = 0
            for r0 in range(t0 * 64, t1 * 64, 16):
                sl = slice(r0, r0 + 16)
                if separate_lo:
                    acc = rnd(acc + X1[sl].T @ R1[sl])
                    acc_lo = rnd(acc_lo + X1[sl].T @ R2[sl])
                    acc_lo = rnd(acc_lo + X2[sl].T @ R1[sl])
                else:
                    acc = rnd(acc + X1[sl].T @ R1[sl])
                    acc = rnd(acc + X1[sl].T @ R2[sl])
                    acc = rnd(acc + X2[sl].T @ R1[sl])
                    if three_way:
                        acc = rnd(acc + X1[sl].T @ R3[sl])
                rows_in_chain += 16
                if rows_in_chain == chain_rows or r0 + 16 == t1 * 64:
                    part = rn32(part + rn32(acc + acc_lo))
                    acc[:] = 0; acc_lo[:] = 0; rows_in_chain = 0
            G += part
        return G

    def resid_of(Z):
        Z = Z.astype(np.float32)
        return (y[:, None] - 1.0 / (1.0 + np.exp(-Z.astype('f8')))).astype(np.float32).astype('f8')

    f32 = (X.T @ (y[:, None] - 1.0 / (1.0 + np.exp(-(X @ W.T)))).astype(np.float32)).astype('f8')
    print('float32 BLAS reference:            elem %.2e norm %.2e' % metrics(f32, G_ref))
    Zt = z_model('rz')
    print('Z error (rz): max %.2e rms %.2e;  rn: max %.2e' % (np.abs(Zt - Z_ref).max(), np.sqrt(np.mean((Zt - Z_ref) ** 2)),
                                                        np.abs(z_model('rn') - Z_ref).max()))
    Z2 = z_model('rz', n_acc=4)
    print('Z error (rz, 4 accumulators over K): max %.2e rms %.2e' % (np.abs(Z2 - Z_ref).max(), np.sqrt(np.mean((Z2 - Z_ref) ** 2))))
    exactG = lambda r: X64.T @ r
    print('only Z truncation (exact G):       elem %.2e norm %.2e' % metrics(exactG(resid_of(Zt)), G_ref))
    print('only Z (4 accs) (exact G):         elem %.2e norm %.2e' % metrics(exactG(resid_of(Z2)), G_ref))
    print('only fp32 rounding of resid:       elem %.2e norm %.2e' % metrics(exactG(resid_of(Z_ref)), G_ref))
    R1, R2 = split(resid_of(Z_ref).astype(np.float32))
    print('only resid bf16 split (exact G):   elem %.2e norm %.2e' % metrics(exactG(R1 + R2), G_ref))
    r0 = resid_of(Z_ref)
    print('only X split, resid exact, no trunc: elem %.2e norm %.2e' % metrics((X1 + X2).T @ r0, G_ref))
    for chain in (2048, 512, 128):
        print('G trunc only (rz, chain %4d):      elem %.2e norm %.2e' % ((chain,) + metrics(g_model(r0, 'rz', chain), G_ref)))
    print('G trunc only (rz, 2048, lo terms in own acc): elem %.2e norm %.2e' % metrics(g_model(r0, 'rz', 2048, separate_lo=True), G_ref))
    print('G rn accumulate (chain 2048):      elem %.2e norm %.2e' % metrics(g_model(r0, 'rn', 2048), G_ref))
    print('G 3-way resid split (rz, 2048):    elem %.2e norm %.2e' % metrics(g_model(r0, 'rz', 2048, three_way=True), G_ref))
    full = g_model(resid_of(Zt), 'rz', 2048)
    print('full model (as the kernel):        elem %.2e norm %.2e' % metrics(full, G_ref))

if __name__ == '__main__':
    main()
