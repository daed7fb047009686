"""The synthetic scaling workload of BASELINE.json configs[4] / SURVEY.md §8(d):
1 arena + 1000 superquadric obstacles x n=100 (10^4 boundary points per surface) x a 32x16 sweep grid,
robot = the rabbit base ellipsoid (10,5,4), orientations uniform on SO(3).  Everything is seeded and
float-rounded like the reference's CSV path (stof), so the oracle and the GPU see identical doubles.
Used by bench.py, __graft_entry__.smoke() and the full-size property tests."""
import numpy as np

from . import hostmath as hm

SEED = 20220726
ARENA = (70.0, 40.0, 30.0)
LIM = [-58.0, 58.0, -28.0, 28.0, -18.0, 18.0]  # arena semi-axes - 1.2 * robot base semi-axis (defineParameters)
BODY_SEMI = [[10.0, 5.0, 4.0]]
N_GRID = 100
LX, LY = 32, 16
N_OBS = 1000
# Obstacle semi-axis ranges.  "scaling" is SURVEY.md §8(d)'s scene (U[2,10]): with the 10 x 5 x 4 robot body its 1000 C-obstacles
# cover nearly the whole arena, so a layer keeps ~0-6 free segments in 512 sweep lines.  "freespace" is the same generator and
# seed with semi-axes U[0.5,2]: the same 1001 surfaces / bytes per layer, but ~45-75 % of the sweep lines carry free segments
# (300-800 segments per layer), which loads the free-segment kernels, the D2H payload and the all-gather.
SEMI_RANGE = {"scaling": (2.0, 10.0), "freespace": (0.5, 2.0)}


def f32round(a):
    return np.asarray(a, dtype=np.float32).astype(np.float64)


def scene_rows12(n_obs=N_OBS, seed=SEED, workload="scaling"):
    """[1+n_obs][12] = a,b,c,e1,e2,px,py,pz,qw,qx,qy,qz (arena first)."""
    rng = np.random.default_rng(seed)
    rows = [[ARENA[0], ARENA[1], ARENA[2], 0.1, 0.1, 0, 0, 0, 1, 0, 0, 0]]
    semi = rng.uniform(SEMI_RANGE[workload][0], SEMI_RANGE[workload][1], size=(n_obs, 3))
    eps = rng.uniform(0.1, 1.8, size=(n_obs, 2))
    pos = rng.uniform(-1, 1, size=(n_obs, 3)) * np.array([60.0, 32.0, 22.0])
    q = rng.normal(size=(n_obs, 4))
    q /= np.linalg.norm(q, axis=1, keepdims=True)
    for i in range(n_obs):
        rows.append(list(semi[i]) + list(eps[i]) + list(pos[i]) + list(q[i]))
    return f32round(np.array(rows, dtype=np.float64))


def scene_rows17(rows12):
    out = np.zeros((rows12.shape[0], 17))
    out[:, :8] = rows12[:, :8]
    for i, r in enumerate(rows12):
        out[i, 8:] = hm.flat9(hm.quat_to_rot(r[8:12]))
    return out


def orientations(k, seed=SEED + 1):
    """k float-rounded quaternions (w,x,y,z), uniform on SO(3)."""
    rng = np.random.default_rng(seed)
    q = rng.normal(size=(k, 4))
    q /= np.linalg.norm(q, axis=1, keepdims=True)
    return f32round(q)


def body_rotations(quats):
    """[K][1][9] per-layer rotation of the single body (per-orientation mode)."""
    return np.array([[hm.flat9(hm.quat_to_rot(q))] for q in quats], dtype=np.float64)


def algorithmic_bytes_per_layer(S, M, L, w=8):
    """C-boundary mesh written once (3*S*M*w) + interval tables (2*L*S*w).
    (CSR segments and inputs are negligible and produced by other kernels.)"""
    return 3 * S * M * w + 2 * L * S * w
