"""The bit-sliced column count of the cluster round kernel's tests (swirld_rcluster.cuh, unit stake): a lane folds the
64-bit masks of its 16 members into a vertical counter with carry-save adders (one 64-bit plane per power of two), the
4 lanes of a test add their counters plane by plane (the shuffles of rc_vadd_xor), and "count > threshold" is evaluated
on the planes from the most significant one down.  Same arithmetic as the kernel, on Python integers."""
import random

M64 = (1 << 64) - 1


def csa(a, b, c):
    """a + b + c = 2 * h + l, per bit column (rc_csa)."""
    u = a ^ b
    return ((a & b) | (u & c)) & M64, (u ^ c) & M64


def lane_counter(x):
    """16 masks -> 5 planes (the CSA tree of the kernel, in its order)."""
    a = [0] * 8
    t2a, a[0] = csa(a[0], x[0], x[1]); t2b, a[0] = csa(a[0], x[2], x[3]); t4a, a[1] = csa(a[1], t2a, t2b)
    t2a, a[0] = csa(a[0], x[4], x[5]); t2b, a[0] = csa(a[0], x[6], x[7]); t4b, a[1] = csa(a[1], t2a, t2b)
    t8a, a[2] = csa(a[2], t4a, t4b)
    t2a, a[0] = csa(a[0], x[8], x[9]); t2b, a[0] = csa(a[0], x[10], x[11]); t4a, a[1] = csa(a[1], t2a, t2b)
    t2a, a[0] = csa(a[0], x[12], x[13]); t2b, a[0] = csa(a[0], x[14], x[15]); t4b, a[1] = csa(a[1], t2a, t2b)
    t8b, a[2] = csa(a[2], t4a, t4b)
    a[4], a[3] = csa(a[3], t8a, t8b)
    return a


def vadd(a, b, nb):
    """nb-plane a + nb-plane b -> nb + 1 planes (rc_vadd_xor: b is the partner lane's counter)."""
    out, carry = list(a), 0
    for k in range(nb):
        carry, out[k] = csa(a[k], b[k], carry)
    out[nb] = carry
    return out


def greater_than(a, thr):
    gt, eq = 0, M64
    for k in range(6, -1, -1):
        tk = M64 if (thr >> k) & 1 else 0
        gt |= eq & a[k] & ~tk & M64
        eq &= ~(a[k] ^ tk) & M64
    return gt


def test_bitsliced_column_counts():
    rng = random.Random(7)
    for trial in range(400):
        dens = rng.random()
        lanes = [[(rng.getrandbits(64) if rng.random() < dens else 0) & (rng.getrandbits(64) | rng.getrandbits(64))
                  for _ in range(16)] for _ in range(4)]
        if trial == 0:
            lanes = [[M64] * 16 for _ in range(4)]          # every member sees every column: counts of 64
        a = [lane_counter(x) for x in lanes]
        b = [vadd(a[l], a[l ^ 1], 5) for l in range(4)]
        c = [vadd(b[l], b[l ^ 2], 6) for l in range(4)]
        thr = rng.randint(0, 63) if trial else 42            # 2 * 64 // 3
        expect = 0
        for col in range(64):
            if sum((m >> col) & 1 for x in lanes for m in x) > thr:
                expect |= 1 << col
        for l in range(4):
            assert greater_than(c[l], thr) == expect
