"""GPU parity of the intra-layer edge tests (K4) and the bridge-layer tests (K5) against the CPU oracle."""
import numpy as np
import pytest

from hrm_b200 import capi
from hrm_b200 import hostmath as hm
from tests import util

pytestmark = pytest.mark.gpu


def layer_vertices3d(sc, off, xM):
    """(tx, ty, xM) of every free segment in (plane, line, segment) order (reference HRM3D.cpp:93-113)."""
    lim, Lx, Ly = sc["lim"], sc["Lx"], sc["Ly"]
    dx = (lim[1] - lim[0]) / float(Lx - 1)
    dy = (lim[3] - lim[2]) / float(Ly - 1)
    v = []
    for ix in range(Lx):
        for iy in range(Ly):
            l = ix * Ly + iy
            for j in range(off[l], off[l + 1]):
                v.append([lim[0] + float(ix) * dx, lim[2] + float(iy) * dy, xM[j]])
    return np.array(v)


@pytest.mark.parametrize("name,n", [("3D_superquadrics_cluttered_rabbit", 10), ("3D_superquadrics_sparse_rabbit", 14)])
def test_edges3d_match_oracle(oracle, gpu_ctx, name, n):
    sc = util.scenes()[name]
    rows = sc["arena"] + sc["obstacle"]
    n_arena, n_obs = len(sc["arena"]), len(sc["obstacle"])
    semi, R, off, quat = util.body_poses3d(sc["robot"], sc["quats"][:3])
    gpu_ctx.set_scene3d(n_arena, n_obs, util.sq17(rows, n), n)
    gpu_ctx.set_sweep(sc["lim"], sc["Lx"], sc["Ly"])
    gpu_ctx.build_layers(semi, R, off, capi.F64_STRICT)
    rng = np.random.default_rng(5)
    B = semi.shape[0]
    all_a, all_b, all_k, all_want = [], [], [], []
    for k in range(3):
        ref = oracle.layer3d(n_arena, n_obs, np.array(rows), n, semi, quat[k], off[k], sc["lim"], sc["Lx"], sc["Ly"])
        o, xL, xU, xM = gpu_ctx.segments(k)
        V = layer_vertices3d(sc, o, xM)
        assert len(V) > 10
        # all vertex pairs at distance <= 3 in the list + vertical detours (bridgeVertex) + random segments through the scene
        i1 = rng.integers(0, len(V), 600)
        i2 = np.clip(i1 + rng.integers(-12, 13, 600), 0, len(V) - 1)
        v1, v2 = V[i1], V[i2]
        det = v1.copy()
        det[:, 2] = v2[:, 2]
        lim = np.array(sc["lim"])
        r1 = rng.uniform(lim[0::2], lim[1::2], size=(300, 3))
        r2 = rng.uniform(lim[0::2], lim[1::2], size=(300, 3))
        a = np.concatenate([v1, v1, det, r1, v1[:5]])
        b = np.concatenate([v2, det, v2, r2, v1[:5]])
        got = gpu_ctx.test_edges(k, a, b)
        want, _ = oracle.edges3d(ref["bound"][n_arena * B:], n, a, b)
        assert np.array_equal(got, want)
        assert 0 < want.sum() < len(want)
        all_a.append(a), all_b.append(b), all_k.append(np.full(len(a), k)), all_want.append(want)
    # the same edges of all layers, interleaved, in ONE launch (hrmgpu_test_edges_multi)
    a, b, kk, want = np.concatenate(all_a), np.concatenate(all_b), np.concatenate(all_k), np.concatenate(all_want)
    perm = rng.permutation(len(a))
    assert np.array_equal(gpu_ctx.test_edges_multi(kk[perm], a[perm], b[perm]), want[perm])
    with pytest.raises(RuntimeError):
        gpu_ctx.test_edges_multi(np.full(len(a), 3), a, b)  # layer index out of range


def test_edges2d_match_oracle(oracle, gpu_ctx):
    sc = util.scenes()["2D_superellipse_cluttered_rabbit"]
    rows = sc["arena"] + sc["obstacle"]
    n_arena, n_obs = len(sc["arena"]), len(sc["obstacle"])
    thetas = [-2.0, 0.3, 1.7]
    semi, ang, off = util.body_poses2d(sc["robot"], thetas)
    gpu_ctx.set_scene2d(n_arena, n_obs, np.array(rows), 50)
    gpu_ctx.set_sweep(sc["lim"], 1, 12)
    gpu_ctx.build_layers2d(semi, ang, off, capi.F64_STRICT)
    rng = np.random.default_rng(11)
    lim = np.array(sc["lim"])
    B = semi.shape[0]
    all_a, all_b, all_k, all_want = [], [], [], []
    for k in range(3):
        ref = oracle.layer2d(n_arena, n_obs, np.array(rows), 50, semi, ang[k], off[k], sc["lim"], 12)
        a = rng.uniform(lim[0::2], lim[1::2], size=(500, 2))
        b = a + rng.normal(scale=8.0, size=(500, 2))
        got = gpu_ctx.test_edges(k, a, b)
        want = oracle.edges2d(ref["bound"][n_arena * B:], a, b)
        assert np.array_equal(got, want)
        assert 0 < want.sum() < len(want)
        all_a.append(a), all_b.append(b), all_k.append(np.full(len(a), k)), all_want.append(want)
    a, b, kk, want = np.concatenate(all_a), np.concatenate(all_b), np.concatenate(all_k), np.concatenate(all_want)
    perm = rng.permutation(len(a))
    assert np.array_equal(gpu_ctx.test_edges_multi(kk[perm], a[perm], b[perm]), want[perm])


def test_bridge3d_match_oracle(oracle, gpu_ctx):
    sc = util.scenes()["3D_superquadrics_cluttered_rabbit"]
    rows = sc["arena"] + sc["obstacle"]
    n_arena, n_obs = len(sc["arena"]), len(sc["obstacle"])
    n = 10
    rb = util.robot3d(sc["robot"])
    semis = [rb.base.semiAxis] + [l.semiAxis for l in rb.link]
    B = len(semis)
    quats = sc["quats"]
    pairs = [(1, 0), (2, 1), (7, 3)]
    tfe_semi, tfe_R, tfe_q = [], [], []
    for (i, j) in pairs:
        ps, pr, pq = [], [], []
        for b in range(B):
            if b == 0:
                qa, qb = quats[i], quats[j]
            else:  # HRM3D.cpp:531-537: Quaterniond(q.toRotationMatrix() * rotLink)
                Rl = rb.tf[b - 1][0]
                qa = hm.rot_to_quat(hm.mat_mul(hm.quat_to_rot(quats[i]), Rl))
                qb = hm.rot_to_quat(hm.mat_mul(hm.quat_to_rot(quats[j]), Rl))
            t = oracle.tfe3d(semis[b], qa, qb, 5)  # TFE itself is host-side work (tiny SVDs)
            ps.append(t[:3])
            pq.append(t[3:])
            pr.append(hm.flat9(hm.quat_to_rot(t[3:])))
        tfe_semi.append(ps), tfe_R.append(pr), tfe_q.append(pq)
    gpu_ctx.set_scene3d(n_arena, n_obs, util.sq17(rows, n), n)
    gpu_ctx.build_bridge(np.array(tfe_semi), np.array(tfe_R), capi.F64_STRICT)
    rng = np.random.default_rng(3)
    lim = np.array(sc["lim"])
    all_c, all_p, all_want = [], [], []
    bounds = [[[oracle.minksum3d(rows[n_arena + o], tfe_semi[p][b], tfe_q[p][b], n, +1) for o in range(n_obs)] for b in range(B)]
              for p in range(len(pairs))]

    def want_for(p, centres):
        want = np.ones(len(centres), dtype=bool)
        for b in range(B):
            w, _ = oracle.pt_in_cfree3d(bounds[p][b], n, centres[:, b, :])
            want &= w
        return want

    for p in range(len(pairs)):
        Q = 400
        centres = rng.uniform(lim[0::2] - 5, lim[1::2] + 5, size=(Q, B, 3))
        got = gpu_ctx.test_bridge(p, centres)
        want = want_for(p, centres)
        assert np.array_equal(got, want)
        assert 0 < want.sum() < Q
        all_c.append(centres), all_p.append(np.full(Q, p)), all_want.append(want)
    # all pairs in ONE launch (hrmgpu_test_bridge_multi)
    c, pp, want = np.concatenate(all_c), np.concatenate(all_p), np.concatenate(all_want)
    perm = rng.permutation(len(c))
    assert np.array_equal(gpu_ctx.test_bridge_multi(pp[perm], c[perm]), want[perm])
    # candidate vertex pairs with the interpolation on the device (hrmgpu_test_transitions): body centres of pose s are
    # w0[s]*v0 + w1[s]*v1 (+ the pair's rotated link translation), a candidate is free iff every pose is
    T, C = 6, 500
    w1 = np.arange(T) * (1.0 / 5.0)
    w0 = 1.0 - w1
    link_off = rng.uniform(-6, 6, size=(len(pairs), T, B, 3))
    pair = rng.integers(0, len(pairs), C)
    v0 = rng.uniform(lim[0::2], lim[1::2], size=(C, 3))
    v1 = v0 + rng.normal(scale=4.0, size=(C, 3))
    got = gpu_ctx.test_transitions(pair, v0, v1, w0, w1, link_off)
    want = np.ones(C, dtype=bool)
    for p in range(len(pairs)):
        sel = np.nonzero(pair == p)[0]
        for s_ in range(T):
            t = w0[s_] * v0[sel] + w1[s_] * v1[sel]
            centres = np.stack([t if b == 0 else link_off[p, s_, b] + t for b in range(B)], axis=1)
            want[sel] &= want_for(p, centres)
    assert np.array_equal(got, want)
    assert 0 < want.sum() < C
    # articulated variant (hrmgpu_test_transitions_articulated): body centre = link_off + (chain_off + t), in that order
    chain_off = rng.uniform(-4, 4, size=(len(pairs), T, B, 3))
    got = gpu_ctx.test_transitions(pair, v0, v1, w0, w1, link_off, chain_off)
    want = np.ones(C, dtype=bool)
    for p in range(len(pairs)):
        sel = np.nonzero(pair == p)[0]
        for s_ in range(T):
            t = w0[s_] * v0[sel] + w1[s_] * v1[sel]
            centres = np.stack([t if b == 0 else link_off[p, s_, b] + (chain_off[p, s_, b] + t) for b in range(B)], axis=1)
            want[sel] &= want_for(p, centres)
    assert np.array_equal(got, want)
    assert 0 < want.sum() < C


def test_bridge2d_match_oracle(oracle, gpu_ctx):
    sc = util.scenes()["2D_superellipse_cluttered_rabbit"]
    rows = sc["arena"] + sc["obstacle"]
    n_arena, n_obs = len(sc["arena"]), len(sc["obstacle"])
    rb = util.robot2d(sc["robot"])
    semis = [rb.base.semiAxis] + [l.semiAxis for l in rb.link]
    B = len(semis)
    tfe = [[oracle.tfe2d(semis[b], -1.0 + 0.2 * b, -0.4 + 0.2 * b, 5) for b in range(B)], [oracle.tfe2d(semis[b], 2.0, 2.6, 5) for b in range(B)]]
    tfe = np.array(tfe)  # [P][B][3] = a, b, angle
    gpu_ctx.set_scene2d(n_arena, n_obs, np.array(rows), 50)
    gpu_ctx.build_bridge2d(tfe[:, :, :2], tfe[:, :, 2], capi.F64_STRICT)
    rng = np.random.default_rng(9)
    lim = np.array(sc["lim"])
    bounds = [[[oracle.minksum2d(rows[n_arena + o], tfe[p, b, :2], tfe[p, b, 2], 50, +1) for o in range(n_obs)] for b in range(B)] for p in range(2)]

    def want_for(p, centres):
        want = np.ones(len(centres), dtype=bool)
        for b in range(B):
            w, _ = oracle.pt_in_cfree2d(bounds[p][b], centres[:, b, :])
            want &= w
        return want

    all_c, all_p, all_want = [], [], []
    for p in range(2):
        Q = 500
        centres = rng.uniform(lim[0::2], lim[1::2], size=(Q, B, 2))
        got = gpu_ctx.test_bridge(p, centres)
        want = want_for(p, centres)
        assert np.array_equal(got, want)
        assert 0 < want.sum() < Q
        all_c.append(centres), all_p.append(np.full(Q, p)), all_want.append(want)
    c, pp, want = np.concatenate(all_c), np.concatenate(all_p), np.concatenate(all_want)
    perm = rng.permutation(len(c))
    assert np.array_equal(gpu_ctx.test_bridge_multi(pp[perm], c[perm]), want[perm])
    # candidate vertex pairs, interpolation on the device (hrmgpu_test_transitions)
    T, C = 5, 600
    w1 = np.arange(T) * (1.0 / 4.0)
    w0 = 1.0 - w1
    link_off = rng.uniform(-4, 4, size=(2, T, B, 2))
    pair = rng.integers(0, 2, C)
    v0 = rng.uniform(lim[0::2], lim[1::2], size=(C, 2))
    v1 = v0 + rng.normal(scale=3.0, size=(C, 2))
    got = gpu_ctx.test_transitions(pair, v0, v1, w0, w1, link_off)
    want = np.ones(C, dtype=bool)
    for p in range(2):
        sel = np.nonzero(pair == p)[0]
        for s_ in range(T):
            t = w0[s_] * v0[sel] + w1[s_] * v1[sel]
            centres = np.stack([t if b == 0 else link_off[p, s_, b] + t for b in range(B)], axis=1)
            want[sel] &= want_for(p, centres)
    assert np.array_equal(got, want)
    assert 0 < want.sum() < C


def test_connect_pairs3d_equals_the_reference_loop(oracle, gpu_ctx):
    """hrmgpu_connect_pairs3d (candidate enumeration from the vertex positions, transition tests and the replay of the moving
    lower bound, all on the device) against the reference's loop (HRM3D.cpp:186-222) written out on the host, with the
    transition tests of hrmgpu_test_transitions (checked against the oracle above)."""
    sc = util.scenes()["3D_superquadrics_cluttered_rabbit"]
    rows = sc["arena"] + sc["obstacle"]
    n_arena, n_obs = len(sc["arena"]), len(sc["obstacle"])
    n = 10
    rb = util.robot3d(sc["robot"])
    semis = [rb.base.semiAxis] + [l.semiAxis for l in rb.link]
    B = len(semis)
    K = 6
    quats = sc["quats"][:K]
    Lx, Ly = 9, 6
    semi, R, off, quat = util.body_poses3d(sc["robot"], quats)
    pairs = [(0, 0), (0, 1), (1, 2), (1, 3), (3, 4), (2, 5)]  # (near, cur); P == number of bridge layers
    tfe_semi, tfe_R = [], []
    for (j, i) in pairs:
        ps, pr = [], []
        for b in range(B):
            if b == 0:
                qa, qb = quats[i], quats[j]
            else:
                Rl = rb.tf[b - 1][0]
                qa = hm.rot_to_quat(hm.mat_mul(hm.quat_to_rot(quats[i]), Rl))
                qb = hm.rot_to_quat(hm.mat_mul(hm.quat_to_rot(quats[j]), Rl))
            t = oracle.tfe3d(semis[b], qa, qb, 5)
            ps.append(t[:3])
            pr.append(hm.flat9(hm.quat_to_rot(t[3:])))
        tfe_semi.append(ps), tfe_R.append(pr)
    gpu_ctx.set_scene3d(n_arena, n_obs, util.sq17(rows, n), n)
    gpu_ctx.set_sweep(sc["lim"], Lx, Ly)
    gpu_ctx.build_layers(semi, R, off, capi.F64_STRICT)
    gpu_ctx.build_bridge(np.array(tfe_semi), np.array(tfe_R), capi.F64_STRICT)
    lim = sc["lim"]
    tx = [lim[0] + float(i) * ((lim[1] - lim[0]) / float(Lx - 1)) for i in range(Lx)]
    ty = [lim[2] + float(i) * ((lim[3] - lim[2]) / float(Ly - 1)) for i in range(Ly)]
    verts = []
    for k in range(K):
        o, xL, xU, xM = gpu_ctx.segments(k)
        verts.append(np.array([[tx[l // Ly], ty[l % Ly], xM[s]] for l in range(Lx * Ly) for s in range(o[l], o[l + 1])]).reshape(-1, 3))
    thx, thy = 2.0 * (lim[1] - lim[0]) / Lx, 2.0 * (lim[3] - lim[2]) / Ly
    T = 6
    w1 = np.arange(T) * (1.0 / 5.0)
    w0 = 1.0 - w1
    rng = np.random.default_rng(11)
    link_off = rng.uniform(-3, 3, size=(len(pairs), T, B, 3))
    # the reference loop on the host: all near candidates tested at once, then the replay
    cand_pair, cand_v0, cand_v1, cand_key = [], [], [], []
    for p, (j, i) in enumerate(pairs):
        for m0, a in enumerate(verts[j]):
            for m1, b in enumerate(verts[i]):
                if abs(a[0] - b[0]) > thx or abs(a[1] - b[1]) > thy:
                    continue
                cand_pair.append(p), cand_v0.append(a), cand_v1.append(b), cand_key.append((p, m0, m1))
    free = gpu_ctx.test_transitions(np.array(cand_pair), np.array(cand_v0), np.array(cand_v1), w0, w1, link_off)
    ok = {key: bool(f) for key, f in zip(cand_key, free)}
    want_of_pair = []
    for p, (j, i) in enumerate(pairs):
        n22 = 0
        want = []
        for m0 in range(len(verts[j])):
            hit = -1
            for m1 in range(n22, len(verts[i])):
                if ok.get((p, m0, m1), False):
                    hit = m1
                    n22 = m1
                    break
            want.append(hit)
        want_of_pair.append(want)
    first = np.concatenate([[0], np.cumsum([len(v) for v in verts])])
    xyz = np.concatenate(verts)
    range4 = [[first[j], first[j + 1], first[i], first[i + 1]] for j, i in pairs]
    got, ncand = gpu_ctx.connect_pairs3d(xyz, range4, thx, thy, w0, w1, link_off)
    want_ids = np.array([first[pairs[p][1]] + w if w >= 0 else -1 for p in range(len(pairs)) for w in want_of_pair[p]])
    assert ncand == len(cand_pair)
    assert np.array_equal(got, want_ids)
    assert (want_ids >= 0).sum() > 20 and (want_ids < 0).sum() > 0
    # error behaviour: P has to be the number of bridge layers
    with pytest.raises(capi.HrmGpuError):
        gpu_ctx.connect_pairs3d(xyz, range4[:1], thx, thy, w0, w1, link_off[:1])


def test_build_bridge_tfe_matches_the_host_ellipsoids(oracle, gpu_ctx):
    """hrmgpu_build_bridge_tfe fits the tightly-fitted ellipsoids of all (pair, body) on the device (TightFitEllipsoid.cpp:43-114):
    semi-axes and orientation against the oracle's to 1e-12 (its libm differs from glibc in the last bits), and the bridge
    layers built from them answer point queries like the ones built from the host ellipsoids."""
    sc = util.scenes()["3D_superquadrics_cluttered_rabbit"]
    rows = sc["arena"] + sc["obstacle"]
    n_arena, n_obs = len(sc["arena"]), len(sc["obstacle"])
    n = 10
    rb = util.robot3d(sc["robot"])
    semis = [rb.base.semiAxis] + [l.semiAxis for l in rb.link]
    B = len(semis)
    quats = sc["quats"]
    pairs = [(1, 0), (2, 1), (7, 3), (30, 12), (59, 58), (5, 5)]
    qa_all, qb_all, want = [], [], []
    for (i, j) in pairs:
        for b in range(B):
            if b == 0:
                qa, qb = quats[i], quats[j]
            else:
                Rl = rb.tf[b - 1][0]
                qa = hm.rot_to_quat(hm.mat_mul(hm.quat_to_rot(quats[i]), Rl))
                qb = hm.rot_to_quat(hm.mat_mul(hm.quat_to_rot(quats[j]), Rl))
            qa_all.append(qa), qb_all.append(qb)
            want.append(oracle.tfe3d(semis[b], qa, qb, 5))
    want = np.array(want).reshape(len(pairs), B, 7)
    gpu_ctx.set_scene3d(n_arena, n_obs, util.sq17(rows, n), n)
    got = gpu_ctx.build_bridge_tfe(np.array(semis), np.array(qa_all), np.array(qb_all), 5, capi.F64_STRICT)
    assert np.max(np.abs(got[..., :3] - want[..., :3]) / want[..., :3]) < 1e-12
    for g, w in zip(got.reshape(-1, 7), want.reshape(-1, 7)):  # the same rotation (q and -q are the same orientation)
        Rg, Rw = np.array(hm.quat_to_rot(g[3:])), np.array(hm.quat_to_rot(w[3:]))
        assert np.max(np.abs(Rg - Rw)) < 1e-10
    # point queries against the device-built bridge layers == against layers built from the host ellipsoids
    rng = np.random.default_rng(5)
    lim = np.array(sc["lim"])
    centres = rng.uniform(lim[0::2] - 5, lim[1::2] + 5, size=(600, B, 3))
    pp = rng.integers(0, len(pairs), 600)
    dev = gpu_ctx.test_bridge_multi(pp, centres)
    gpu_ctx.build_bridge(want[..., :3], np.array([[hm.flat9(hm.quat_to_rot(w[3:])) for w in pr] for pr in want]), capi.F64_STRICT)
    host = gpu_ctx.test_bridge_multi(pp, centres)
    assert np.array_equal(dev, host) and 0 < host.sum() < 600


def test_connect_slices3d_equals_host_enumeration(gpu_ctx):
    """hrmgpu_connect_slices3d (candidate segments of connectOneSlice generated on the device from the free segments, all layers in
    one call) against the same segments enumerated on the host in the reference's order and tested with hrmgpu_test_edges_multi."""
    sc = util.scenes()["3D_superquadrics_cluttered_rabbit"]
    rows = sc["arena"] + sc["obstacle"]
    n = 10
    K, Lx, Ly = 5, 8, 5
    semi, R, off, quat = util.body_poses3d(sc["robot"], sc["quats"][:K])
    gpu_ctx.set_scene3d(len(sc["arena"]), len(sc["obstacle"]), util.sq17(rows, n), n)
    gpu_ctx.set_sweep(sc["lim"], Lx, Ly)
    gpu_ctx.build_layers(semi, R, off, capi.F64_STRICT)
    lim = sc["lim"]
    tx = [lim[0] + float(i) * ((lim[1] - lim[0]) / float(Lx - 1)) for i in range(Lx)]
    ty = [lim[2] + float(i) * ((lim[3] - lim[2]) / float(Ly - 1)) for i in range(Ly)]
    A, Bq, layer, first = [], [], [], [0]
    for k in range(K):
        o, xL, xU, xM = gpu_ctx.segments(k)
        pairs = []
        for ix in range(Lx):
            for iy in range(Ly - 1):
                for j1 in range(o[ix * Ly + iy], o[ix * Ly + iy + 1]):
                    for j2 in range(o[ix * Ly + iy + 1], o[ix * Ly + iy + 2]):
                        pairs.append(((tx[ix], ty[iy], xM[j1]), (tx[ix], ty[iy + 1], xM[j2])))
        for ix in range(Lx - 1):
            for iy in range(Ly):
                for k1 in range(o[ix * Ly + iy], o[ix * Ly + iy + 1]):
                    for k2 in range(o[(ix + 1) * Ly + iy], o[(ix + 1) * Ly + iy + 1]):
                        pairs.append(((tx[ix], ty[iy], xM[k1]), (tx[ix + 1], ty[iy], xM[k2])))
        for v1, v2 in pairs:
            w1, w2 = (v1[0], v1[1], v2[2]), (v2[0], v2[1], v1[2])
            A += [v1, v1, w1, v1, w2]
            Bq += [v2, w1, v2, w2, v2]
            layer += [k] * 5
        first.append(first[-1] + len(pairs))
    f = gpu_ctx.test_edges_multi(np.array(layer), np.array(A), np.array(Bq)).reshape(-1, 5)
    want = np.where(f[:, 0], 0, np.where(f[:, 1] & f[:, 2], 1, np.where(f[:, 3] & f[:, 4], 2, 3)))
    got, got_first = gpu_ctx.connect_slices3d(first[-1])
    assert np.array_equal(got_first, np.array(first))
    assert np.array_equal(got, want)
    assert len(set(want.tolist())) >= 3  # direct connections, detours and failures all occur
    with pytest.raises(capi.HrmGpuError) as e:  # answer buffer too small
        gpu_ctx.connect_slices3d(first[-1] - 1)
    assert e.value.code == capi.E_CAP


@pytest.mark.parametrize("precision", [capi.F64_STRICT, capi.F64, capi.F32])
def test_tabulated_boxes_and_normals_change_no_answer(monkeypatch, precision):
    """The planner-side kernels with their per-mesh tables -- band / face xy boxes and the queued line-triangle tests of
    k_transitions3d, unit face normals of k_edges3d -- against the same kernels without them (HRMGPU_NO_BANDS /
    HRMGPU_NO_FACE_NORMALS: every face walked, every normal recomputed), on crowded inputs: 20 000 candidate pairs and 30 000
    segments concentrated around the obstacles, so that warps hold many surviving items and the queue drains mid-item."""
    sc = util.scenes()["3D_superquadrics_cluttered_rabbit"]
    rows = sc["arena"] + sc["obstacle"]
    n_arena, n_obs = len(sc["arena"]), len(sc["obstacle"])
    n = 10
    K = 4
    semi, R, off, quat = util.body_poses3d(sc["robot"], sc["quats"][:K])
    B = semi.shape[0]
    rng = np.random.default_rng(11)
    obs = np.array(sc["obstacle"])
    centre = obs[rng.integers(0, n_obs, 30000), 5:8]  # (rows: semi-axes 0-2, eps 3-4, position 5-7, quaternion 8-11)
    a = centre + rng.normal(scale=6.0, size=centre.shape)
    b = a + rng.normal(scale=5.0, size=centre.shape)
    axis = rng.integers(0, 4, len(a))  # a share of exactly axis-parallel segments, like the detours of bridgeVertex
    for d in range(3):
        keep = (axis != 3) & (axis != d)
        b[keep, d] = a[keep, d]
    layer = rng.integers(0, K, len(a))
    P, T, C = 3, 6, 20000
    tfe_semi = rng.uniform(3.0, 9.0, size=(P, B, 3))
    tfe_R = np.array([[hm.flat9(hm.quat_to_rot((q / np.linalg.norm(q)).tolist())) for q in rng.normal(size=(B, 4))] for _ in range(P)])
    w1 = np.arange(T) * (1.0 / 5.0)
    w0 = 1.0 - w1
    link_off = rng.uniform(-6, 6, size=(P, T, B, 3))
    pair = rng.integers(0, P, C)
    v0 = obs[rng.integers(0, n_obs, C), 5:8] + rng.normal(scale=9.0, size=(C, 3))
    v1 = v0 + rng.normal(scale=3.0, size=(C, 3))
    out = {}
    for plain in ("0", "1"):
        monkeypatch.setenv("HRMGPU_NO_BANDS", plain)
        monkeypatch.setenv("HRMGPU_NO_FACE_NORMALS", plain)
        ctx = capi.Context(0)
        try:
            ctx.set_scene3d(n_arena, n_obs, util.sq17(rows, n), n)
            ctx.set_sweep(sc["lim"], sc["Lx"], sc["Ly"])
            ctx.build_layers(semi, R, off, precision)
            ctx.build_bridge(tfe_semi, tfe_R, precision)
            out[plain] = (ctx.test_edges_multi(layer, a, b), ctx.test_transitions(pair, v0, v1, w0, w1, link_off))
        finally:
            ctx.close()
    for got, want in zip(out["0"], out["1"]):
        assert np.array_equal(got, want)
        assert 0.05 < want.mean() < 0.95


@pytest.mark.parametrize("n", [33, 36])
def test_transitions_on_fine_meshes_match_oracle(oracle, gpu_ctx, n):
    """n = 33 is the finest mesh the band / face-box tables cover (32 bands per direction, face indices up to 2047 in the queue
    entries); n = 36 takes the plain face walk.  Both against the oracle's point-in-C-free test."""
    sc = util.scenes()["3D_superquadrics_cluttered_rabbit"]
    rows = sc["arena"] + sc["obstacle"]
    n_arena, n_obs = len(sc["arena"]), len(sc["obstacle"])
    rng = np.random.default_rng(17)
    B, T, C = 2, 4, 160
    tfe_semi = rng.uniform(3.0, 9.0, size=(1, B, 3))
    quat = [(q / np.linalg.norm(q)).tolist() for q in rng.normal(size=(B, 4))]
    tfe_R = np.array([[hm.flat9(hm.quat_to_rot(q)) for q in quat]])
    gpu_ctx.set_scene3d(n_arena, n_obs, util.sq17(rows, n), n)
    gpu_ctx.build_bridge(tfe_semi, tfe_R, capi.F64_STRICT)
    bounds = [[oracle.minksum3d(rows[n_arena + o], tfe_semi[0][b], quat[b], n, +1) for o in range(n_obs)] for b in range(B)]
    w1 = np.arange(T) * (1.0 / (T - 1))
    w0 = 1.0 - w1
    link_off = rng.uniform(-6, 6, size=(1, T, B, 3))
    obs = np.array(sc["obstacle"])
    v0 = obs[rng.integers(0, n_obs, C), 5:8] + rng.normal(scale=12.0, size=(C, 3))
    v1 = v0 + rng.normal(scale=3.0, size=(C, 3))
    got = gpu_ctx.test_transitions(np.zeros(C, dtype=np.int64), v0, v1, w0, w1, link_off)
    want = np.ones(C, dtype=bool)
    for s_ in range(T):
        t = w0[s_] * v0 + w1[s_] * v1
        for b in range(B):
            centres = t if b == 0 else link_off[0, s_, b] + t
            w, _ = oracle.pt_in_cfree3d(bounds[b], n, centres)
            want &= w
    assert np.array_equal(got, want)
    assert 0.1 < want.mean() < 0.9
