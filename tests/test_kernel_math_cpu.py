"""CPU checks of the arithmetic the packed SU(2) kernel relies on (csrc/uqoc_su2_x2.cuh), restated in numpy:

* the generated sin/cos tables (csrc/uqoc_sincos_table.inc) and the three-instruction table-index arithmetic on the
  shifted-node tables (``UQOC_X2_IDX3``): one rounding in ``kc``, centred residuals, exact identity for padding rows;
* the two identities of the backward core (``UQOC_X2_BWD_CORE = 1``): d/dphi as the telescoping z-torque and the
  invariant axis component, against the direct form and against the oracle's gradient.

No GPU, no library call: the kernels themselves are checked against the oracle in the ``-m gpu`` tests."""
import os
import re
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
INC = os.path.join(ROOT, "universal_quantum_optimal_control_b200", "csrc", "uqoc_sincos_table.inc")
f32 = np.float32


def _inc():
    txt = open(INC).read()
    defs = dict(re.findall(r"#define (UQOC_SINCOS_\w+) (\S+)", txt))

    def arr(name):
        body = re.search(r"%s\[\d+\] = \{(.*?)\};" % re.escape(name), txt, re.S).group(1)
        return np.array([float.fromhex(v.rstrip("f")) for v in re.findall(r"-?0x[0-9a-fp.+-]+f?", body)])

    def num(key):
        v = defs[key].rstrip("f")
        return float.fromhex(v) if "x" in v else float(v)

    return defs, arr, num


def fma32(a, b, c):
    """float32 FMA: products of two float32 are exact in double; the sum is rounded to double, then to float32 (a double
    rounding that can differ from a true FMA only at exact ties of the float32 rounding - irrelevant for the bounds here)."""
    return (np.asarray(a, np.float64) * np.asarray(b, np.float64) + np.asarray(c, np.float64)).astype(f32)


def test_shifted_node_tables_and_constants():
    defs, arr, num = _inc()
    N = int(defs["UQOC_SINCOS_TABLE_N"])
    M3, step, C0, e0 = (num("UQOC_SINCOS_" + k) for k in ("M3", "STEP", "C0", "E0"))
    assert N == 1024 and int(defs["UQOC_SINCOS_TABLE_LEN"]) == 2 * N
    # the rounding constant: an integer-valued float in [2^23, 2^24) (ulp 1) and a multiple of the table period
    assert 2 ** 23 <= M3 < 2 ** 24 and M3 == int(M3) and int(M3) % (2 * N) == 0 and float(f32(M3)) == M3
    assert step == float(f32(np.pi / N))
    # C0 = float(step * M3); e0 is what the one rounding leaves: exact, tiny, a float
    assert C0 == float(f32(step * M3)) and e0 == step * M3 - C0 and float(f32(e0)) == e0
    assert abs(e0) < 1e-3 * step                                       # the residuals stay centred on the nodes
    sin_sh, cos_sh, il_sh = arr("g_sin_table_sh"), arr("g_cos_table_sh"), arr("g_sincos_table_sh")
    sin0, cos0 = arr("g_sin_table"), arr("g_cos_table")
    assert len(sin_sh) == len(cos_sh) == 2 * N and len(il_sh) == 4 * N
    assert np.array_equal(il_sh[0::2], sin_sh) and np.array_equal(il_sh[1::2], cos_sh)
    k = np.arange(2 * N)
    scale = 1.0 - (np.pi / (2 * N)) ** 2 / 6.0
    inner = np.ones(2 * N, bool)
    inner[[0, N]] = False
    assert np.abs(sin_sh - np.sin(k * np.pi / N + e0) * scale)[inner].max() < 1e-7
    assert np.abs(cos_sh - np.cos(k * np.pi / N + e0) * scale)[inner].max() < 1e-7
    assert np.abs(sin0 - np.sin(k * np.pi / N) * scale)[inner].max() < 1e-7 and np.abs(cos0 - np.cos(k * np.pi / N) * scale)[inner].max() < 1e-7
    # period: entry k + N is minus entry k (the half-period users drop that sign)
    assert np.array_equal(sin_sh[N:], -sin_sh[:N]) or np.abs(sin_sh[N:] + sin_sh[:N]).max() < 1.2e-7
    assert np.abs(cos_sh[N:] + cos_sh[:N]).max() < 1.2e-7
    # a zero-duration padding row (tau = 0: kf = M3, kc = e0, r = -e0) is the EXACT identity on entry 0
    kc = fma32(f32(M3), f32(step), f32(-C0))
    assert float(kc) == e0
    r = fma32(f32(0.0), f32(1.2345), -kc)
    st, ct = f32(sin_sh[0]), f32(cos_sh[0])
    assert float(fma32(r, ct, st)) == 0.0 and float(fma32(-r, st, ct)) == 1.0
    st, ct = f32(sin_sh[N]), f32(cos_sh[N])
    assert float(fma32(r, ct, st)) == 0.0 and float(fma32(-r, st, ct)) == -1.0


def test_three_instruction_index_arithmetic_matches_the_exact_residual():
    """kf = fma(tau, a N/pi, M3); kc = fma(kf, pi/N, -C0); r = fma(tau, a, -kc): k in the low mantissa bits, r the residual
    from the node k pi/N + e0 - within half an ulp of the angle of the exact value, for negative and multi-turn angles too."""
    _, arr, num = _inc()
    N = 1024
    M3, step, C0, e0 = (num("UQOC_SINCOS_" + k) for k in ("M3", "STEP", "C0", "E0"))
    sin_sh, cos_sh = arr("g_sin_table_sh"), arr("g_cos_table_sh")
    rng = np.random.default_rng(0)
    for amax in (0.7, 8.0, 60.0):
        a64 = rng.uniform(0.5, 3.0, 20000)                              # slope (1 + eps) w, in double as make_sample_const has it
        tau = (rng.uniform(-1, 1, 20000) * amax / 3.0).astype(f32)
        a, kap = a64.astype(f32), (a64 * (N / np.pi)).astype(f32)       # both rounded once from double
        kf = fma32(tau, kap, f32(M3))
        k = (kf.view(np.int32) & 0x7FFFFF).astype(np.int64) - (int(M3) - 2 ** 23)      # low mantissa bits, M3's removed
        assert np.array_equal(k, np.rint(tau.astype(np.float64) * kap.astype(np.float64)).astype(np.int64)) or \
            np.abs(k - tau.astype(np.float64) * kap.astype(np.float64)).max() <= 0.5 + 1e-6
        kc = fma32(kf, f32(step), f32(-C0))
        r = fma32(tau, a, -kc)
        angle = tau.astype(np.float64) * a.astype(np.float64)           # the angle the float slope stands for
        r_exact = angle - (k * np.pi / N + e0)
        ulp = np.spacing(np.maximum(np.abs(angle), np.abs(kc.astype(np.float64))).astype(f32)).astype(np.float64)
        # one rounding of kc (half an ulp of the angle) + the table step constant being float(pi/N) (2.8e-8 relative) + the
        # final rounding of r itself
        assert np.all(np.abs(r - r_exact) <= 0.5 * ulp + 3e-8 * np.abs(angle) + np.spacing(np.abs(r)) + 1e-12)
        assert np.abs(r).max() <= 0.5 * step * 1.01 + abs(e0) + 4e-6 * amax
        # and the first-order table step reproduces sin / cos of the full angle (full-period table, index k mod 2N)
        idx = np.mod(k, 2 * N)
        st, ct = sin_sh[idx].astype(f32), cos_sh[idx].astype(f32)
        s, c = fma32(r, ct, st), fma32(-r, st, ct)
        tol = 1.6e-6 + 0.6 * ulp.max() + 3e-8 * amax                     # radial error r^2 / 2 of the first-order step + the angle's
        assert np.abs(s - np.sin(angle)).max() < tol and np.abs(c - np.cos(angle)).max() < tol
        # the angle of (s, c) - what the rotation actually uses: float rounding of the table entries and of (s, c) themselves
        # (~1e-7), the first-order step's r^3 / 3 (1.2e-9), the rounding of kc
        dang = np.angle((c.astype(np.float64) + 1j * s.astype(np.float64)) * np.exp(-1j * angle))
        assert np.abs(dang).max() < 2e-7 + 0.6 * ulp.max() + 3e-8 * amax


def _backward_reference(phi, tau, delta, eps, lam):
    """Adjoint sweep of ONE sample in double, direct form: returns per-pulse (dphi, dtau) of
    <lam, P_L> for the quaternion product P_L of the pulses, and the W3 sequence entering each pulse."""
    L = len(phi)
    w = np.sqrt(1 + delta * delta)
    a, r = 0.5 * (1 + eps) * w, 1 / w

    def qmul(p, q):
        return np.array([p[0] * q[0] - p[1] * q[1] - p[2] * q[2] - p[3] * q[3], p[0] * q[1] + p[1] * q[0] + p[2] * q[3] - p[3] * q[2],
                         p[0] * q[2] - p[1] * q[3] + p[2] * q[0] + p[3] * q[1], p[0] * q[3] + p[1] * q[2] - p[2] * q[1] + p[3] * q[0]])

    P = np.array([1.0, 0, 0, 0])
    for i in range(L):
        h = tau[i] * a
        q = np.array([np.cos(h), r * np.sin(h) * np.cos(phi[i]), r * np.sin(h) * np.sin(phi[i]), r * np.sin(h) * delta])
        P = qmul(q, P)
    Wq = qmul(lam, P * np.array([1, -1, -1, -1]))
    A = Wq[1] * np.cos(phi[-1]) + Wq[2] * np.sin(phi[-1])
    B = Wq[2] * np.cos(phi[-1]) - Wq[1] * np.sin(phi[-1])
    W3 = Wq[3]
    gphi, gtau, W3_in = np.zeros(L), np.zeros(L), np.zeros(L + 1)
    for i in range(L - 1, -1, -1):
        W3_in[i + 1] = W3
        C2, S2 = np.cos(2 * tau[i] * a), np.sin(2 * tau[i] * a)
        Sr, k1 = r * S2, r * r * (1 - C2)
        t, uu = A + delta * W3, delta * A - W3
        gtau[i] = 0.5 * (1 + eps) * t
        gphi[i] = 0.5 * (Sr * B - k1 * uu)
        K, BS = k1 * t, B * Sr
        A1 = A * C2 + K + delta * BS
        B1 = B * C2 - uu * Sr
        W3 = W3 * C2 - BS + delta * K
        assert abs((t - delta * W3) - A1) < 1e-13 * (1 + abs(A1))          # identity (b): the axis component is invariant
        dphi = phi[i] - phi[i - 1] if i > 0 else 0.0
        A, B = A1 * np.cos(dphi) - B1 * np.sin(dphi), A1 * np.sin(dphi) + B1 * np.cos(dphi)
    W3_in[0] = W3
    return gphi, gtau, W3_in, P


def test_backward_core_identities_against_direct_form_and_finite_differences():
    rng = np.random.default_rng(3)
    L = 23
    phi, tau = rng.uniform(-3.1, 3.1, L), rng.uniform(0.05, 0.6, L)
    lam = rng.normal(size=4)
    for delta, eps in ((0.0, 0.0), (0.8, -0.04), (-2.5, 0.11)):
        gphi, gtau, W3_in, P = _backward_reference(phi, tau, delta, eps, lam)
        # identity (a): d/dphi_i = (W3 entering pulse i - W3 leaving it) / 2, and W3 is untouched by the frame change, so the
        # sequence telescopes: W3_in[i + 1] enters pulse i, W3_in[i] leaves it
        assert np.abs(gphi - 0.5 * (W3_in[1:] - W3_in[:-1])).max() < 1e-13
        # both against central finite differences of <lam, P_L>
        def f(ph, ta):
            return float(np.dot(lam, _backward_reference(ph, ta, delta, eps, lam)[3]))
        for i in (0, 7, L - 1):
            d = 1e-6
            e = np.zeros(L)
            e[i] = d
            assert abs((f(phi + e, tau) - f(phi - e, tau)) / (2 * d) - gphi[i]) < 1e-8
            assert abs((f(phi, tau + e) - f(phi, tau - e)) / (2 * d) - gtau[i]) < 1e-8
    # summed over samples the telescoped form needs only S_i = sum_s W3_s: the differences of the sums are the summed gradients
    tot, S = np.zeros(L), np.zeros(L + 1)
    for s in range(50):
        g, _, w3, _ = _backward_reference(phi, tau, rng.normal(), 0.05 * rng.normal(), lam)
        tot += g
        S += w3
    assert np.abs(tot - 0.5 * (S[1:] - S[:-1])).max() < 1e-12
