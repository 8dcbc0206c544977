"""Host logic of the stream-K schedule of csrc/conv2d_tc.cu (no GPU): the partition of the tile-major (tile, K-step) list into one
contiguous range per cluster, through the library's own rule (dvd_conv2d_streamk_bounds = SegIter::bound on the host).

Invariants the kernel's exchange protocol relies on:
  * ranges are monotone, cover every K-step of every tile exactly once, and range boundaries within 1/8 of a tile boundary snap to it;
  * a cluster delivers at most ONE partial tile (the first segment of its range, when that range starts inside a tile);
  * the cluster that owns the first K-steps of a split tile finds exactly the clusters whose ranges BEGIN inside that tile as its
    contributors (the kernel counts them with the same rule), so every raised flag is consumed and lowered in the same launch."""
import ctypes
import random

import pytest


def _bounds(ntiles, ksteps, n):
    from dvd_b200 import _lib
    out = (ctypes.c_long * (n + 1))()
    _lib.check(_lib.load().dvd_conv2d_streamk_bounds(ntiles, ksteps, n, out), 'bounds')
    return list(out)


def _check(ntiles, ksteps, n):
    b = _bounds(ntiles, ksteps, n)
    total = ntiles * ksteps
    assert b[0] == 0 and b[-1] == total and all(b[i] <= b[i + 1] for i in range(n))
    for u in b:
        r = u % ksteps
        assert r == 0 or (r * 8 >= ksteps and (ksteps - r) * 8 >= ksteps), (u, ksteps)       # slivers are snapped away
    flags, owners, cover = {}, [], {}
    for c in range(n):
        u, u1 = b[c], b[c + 1]
        first = True
        while u < u1:
            t, k0 = divmod(u, ksteps)
            k1 = min(ksteps, k0 + (u1 - u))
            cover.setdefault(t, []).append((k0, k1))
            if k0 > 0:
                assert first and c not in flags     # only the first segment of a range can be a partial tile
                flags[c] = t
            elif k1 < ksteps:
                npart = 0
                while c + 1 + npart < n and b[c + 1 + npart] < (t + 1) * ksteps:
                    npart += 1
                owners.append((c, t, npart))
            u += k1 - k0
            first = False
    for c, t, npart in owners:
        assert npart >= 1
        for p in range(npart):
            assert flags.pop(c + 1 + p) == t
    assert not flags
    for t in range(ntiles):
        pos = 0
        for k0, k1 in sorted(cover[t]):
            assert k0 == pos
            pos = k1
        assert pos == ksteps


@pytest.mark.parametrize('shape', [(84, 32, 74), (6, 576, 74), (24, 288, 74), (1, 64, 4), (48, 64, 74), (168, 16, 74), (11, 8, 11), (3, 9, 2)])
def test_partition_of_the_layers_of_the_depth_net(shape):
    _check(*shape)


def test_partition_invariants_on_random_shapes():
    rnd = random.Random(0)
    done = 0
    while done < 3000:
        ks, n, nt = rnd.randint(1, 130), rnd.randint(1, 74), rnd.randint(1, 400)
        if nt * ks < n:
            continue
        _check(nt, ks, n)
        done += 1
