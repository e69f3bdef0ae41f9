"""The lower-envelope column pass of the exact EDT (edt_env_kernel in voxlogica_b200/csrc/vx_edt.cu)
restated on Python integers: a stack of parabolas (position s, first index t where it is the
minimum, height g), popped while the new parabola beats the top at the start of its interval, the
separation index from a binary32 estimate corrected in integers, and a read-out sweep.  Checked
against the brute-force minimum  out[i] = min_j g[j] + s2 * (i - j)^2  (integer spacings: exact)."""
import numpy as np
import pytest

INF = None


def env_pass(g, s2):
    n = len(g)
    st = []                                    # entries (s, t, gval)
    for u in range(n):
        if g[u] is INF:
            continue
        gu = int(g[u])
        if not st:
            st.append((u, 0, gu))
            continue
        only = False
        while True:
            ts, tt, tg = st[-1]
            d1, d2 = tt - ts, tt - u
            if tg + s2 * d1 * d1 > gu + s2 * d2 * d2:
                if len(st) == 1:
                    only = True
                    break
                st.pop()
            else:
                break
        if only:
            st[-1] = (u, 0, gu)
            continue
        ts, tt, tg = st[-1]
        num, den = s2 * (u * u - ts * ts) + gu - tg, 2 * s2 * (u - ts)
        sep = int(np.floor(np.float32(num) / np.float32(den)))     # the kernel's binary32 estimate ...
        while sep * den > num:                                      # ... and its exact integer correction
            sep -= 1
        while (sep + 1) * den <= num:
            sep += 1
        w = sep + 1
        if w < n:
            st.append((u, w, gu))
    out = [INF] * n
    if st:
        k = 0
        for u in range(n):
            while k + 1 < len(st) and st[k + 1][1] == u:
                k += 1
            s, _, gv = st[k]
            out[u] = gv + s2 * (u - s) * (u - s)
    return out


def brute(g, s2):
    n = len(g)
    out = []
    for i in range(n):
        c = [g[j] + s2 * (i - j) ** 2 for j in range(n) if g[j] is not INF]
        out.append(min(c) if c else INF)
    return out


@pytest.mark.parametrize("n", [1, 2, 7, 33, 120])
@pytest.mark.parametrize("s2", [1, 4, 9])
@pytest.mark.parametrize("density", [0.05, 0.4, 1.0])
def test_envelope_pass_is_the_exact_min_plus_transform(n, s2, density):
    rng = np.random.default_rng(n * 100 + s2 * 10 + int(density * 10))
    for trial in range(6):
        g = [int(v) if rng.random() < density else INF for v in rng.integers(0, 4000 if trial % 2 else 30, n)]
        assert env_pass(g, s2) == brute(g, s2)
    assert env_pass([INF] * n, s2) == [INF] * n
    # plateaus and equal heights: ties between parabolas must still give the minimum
    g = [5] * n
    assert env_pass(g, s2) == brute(g, s2)
