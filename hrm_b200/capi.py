"""ctypes binding of libhrmgpu.so (include/hrmgpu.h).  Loading fails loudly if the CUDA library has not
been built; there is no CPU fallback anywhere in this package."""
import ctypes as C
import os

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "lib", "libhrmgpu.so")

OK, E_ARG, E_CUDA, E_STATE, E_CAP, E_NOMEM, E_NODEVICE = 0, -1, -2, -3, -4, -5, -6
F64_STRICT, F64, F32 = 0, 1, 2

_STATUS = {E_ARG: "E_ARG", E_CUDA: "E_CUDA", E_STATE: "E_STATE", E_CAP: "E_CAP", E_NOMEM: "E_NOMEM", E_NODEVICE: "E_NODEVICE"}

_lib = None


class HrmGpuError(RuntimeError):
    def __init__(self, code, msg):
        super().__init__("hrmgpu %s (%d): %s" % (_STATUS.get(code, "?"), code, msg))
        self.code = code


def load():
    """Loads libhrmgpu.so; raises if it is missing (build with `python -m hrm_b200.build`)."""
    global _lib
    if _lib is not None:
        return _lib
    path = os.environ.get("HRMGPU_LIB", LIB_PATH)  # (experiment builds of the same library: tools/build_variant.sh)
    if not os.path.exists(path):
        raise RuntimeError("libhrmgpu.so not built: run `python -m hrm_b200.build` (there is no CPU fallback)")
    lib = C.CDLL(path)
    vp, i, ll = C.c_void_p, C.c_int, C.c_longlong
    sig = {
        "hrmgpu_abi_version": (i, []),
        "hrmgpu_device_count": (i, []),
        "hrmgpu_create": (i, [C.POINTER(vp), i, C.c_uint]),
        "hrmgpu_destroy": (None, [vp]),
        "hrmgpu_reset": (i, [vp]),
        "hrmgpu_last_error": (C.c_char_p, [vp]),
        "hrmgpu_set_stream": (i, [vp, vp]),
        "hrmgpu_synchronize": (i, [vp]),
        "hrmgpu_set_scene3d": (i, [vp, i, i, vp, i]),
        "hrmgpu_set_scene2d": (i, [vp, i, i, vp, i]),
        "hrmgpu_set_sweep": (i, [vp, vp, i, i]),
        "hrmgpu_resweep": (i, [vp, vp, i, i]),
        "hrmgpu_build_layers": (i, [vp, i, i, vp, vp, vp, i]),
        "hrmgpu_build_layers_dev": (i, [vp, i, i, vp, vp, vp, i]),
        "hrmgpu_build_layers2d": (i, [vp, i, i, vp, vp, vp, i]),
        "hrmgpu_get_dims": (i, [vp, vp]),
        "hrmgpu_get_boundary": (i, [vp, i, i, vp]),
        "hrmgpu_get_intervals": (i, [vp, i, vp, vp]),
        "hrmgpu_get_segments": (i, [vp, i, vp, vp, vp, vp, i]),
        "hrmgpu_get_segments_all": (i, [vp, vp, vp, vp, vp, ll]),
        "hrmgpu_get_single_hit_count": (i, [vp, vp]),
        "hrmgpu_test_edges": (i, [vp, i, i, vp, vp, vp]),
        "hrmgpu_test_edges_multi": (i, [vp, i, vp, vp, vp, vp]),
        "hrmgpu_test_bridge_multi": (i, [vp, i, vp, vp, vp]),
        "hrmgpu_test_transitions": (i, [vp, i, vp, vp, vp, i, vp, vp, vp, vp]),
        "hrmgpu_test_transitions_articulated": (i, [vp, i, vp, vp, vp, i, vp, vp, vp, vp, vp]),
        "hrmgpu_connect_slices3d": (ll, [vp, vp, ll, vp]),
        "hrmgpu_connect_pairs3d": (ll, [vp, ll, vp, i, vp, C.c_double, C.c_double, i, vp, vp, vp, vp, vp, ll, vp]),
        "hrmgpu_build_bridge": (i, [vp, i, i, vp, vp, i]),
        "hrmgpu_build_bridge_tfe": (i, [vp, i, i, vp, vp, vp, i, i, vp]),
        "hrmgpu_build_bridge2d": (i, [vp, i, i, vp, vp, i]),
        "hrmgpu_test_bridge": (i, [vp, i, i, vp, vp]),
        "hrmgpu_pack_segments_dev": (ll, [vp, vp, ll]),
        "hrmgpu_join": (i, [vp]),
        "hrmgpu_get_boundary_layer": (i, [vp, i, vp]),
        "hrmgpu_fetch_segments_async": (ll, [vp, vp, vp, vp, vp, ll]),
        "hrmgpu_wait_segments": (ll, [vp]),
        "hrmgpu_comm_unique_id": (i, [vp]),
        "hrmgpu_comm_init": (i, [vp, i, i, vp]),
        "hrmgpu_comm_destroy": (i, [vp]),
        "hrmgpu_allgather_layers": (i, [vp, vp, vp, ll, ll]),
        "hrmgpu_gathered_info": (i, [vp, vp, vp]),
        "hrmgpu_get_gathered": (ll, [vp, i, vp, vp, vp, vp, vp, ll]),
        "hrmgpu_last_segments_ms": (C.c_double, [vp]),
        "hrmgpu_segment_redos": (ll, [vp, i]),
        "hrmgpu_kernel_launches": (ll, [vp, i]),
        "hrmgpu_last_minksweep_ms": (C.c_double, [vp]),
    }
    for name, (res, args) in sig.items():
        fn = getattr(lib, name)  # AttributeError if the library does not export a declared symbol
        fn.restype = res
        fn.argtypes = args
    _lib = lib
    return lib


def declared_symbols():
    """Names of every function include/hrmgpu.h declares (parsed from the header)."""
    import re

    hdr = os.path.join(_HERE, "..", "include", "hrmgpu.h")
    txt = open(hdr).read()
    return sorted(set(re.findall(r"\b(hrmgpu_[a-z0-9_]+)\s*\(", txt)))


def _d(a):
    return np.ascontiguousarray(a, dtype=np.float64)


def _p(a):
    return None if a is None else a.ctypes.data_as(C.c_void_p)


class Context:
    """One hrmgpu context (one GPU).  Thin, argument-checked wrapper; arrays are numpy float64."""

    def __init__(self, device=0):
        self.lib = load()
        h = C.c_void_p()
        rc = self.lib.hrmgpu_create(C.byref(h), device, 0)
        if rc != OK:
            raise HrmGpuError(rc, "hrmgpu_create failed (no CUDA device? there is no CPU fallback)")
        self.h = h

    def close(self):
        if getattr(self, "h", None):
            self.lib.hrmgpu_destroy(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def _chk(self, rc):
        if rc < 0:
            raise HrmGpuError(rc, self.lib.hrmgpu_last_error(self.h).decode())
        return rc

    # ---- setup ----
    def last_error(self):
        """Message of the last failed call (hrmgpu_last_error)."""
        return self.lib.hrmgpu_last_error(self.h).decode()

    def set_stream(self, cuda_stream_ptr):
        self._chk(self.lib.hrmgpu_set_stream(self.h, C.c_void_p(cuda_stream_ptr)))

    def synchronize(self):
        self._chk(self.lib.hrmgpu_synchronize(self.h))

    def join(self):
        """Main stream waits (on the device) for the side stream: free-segment passes, exchange."""
        self._chk(self.lib.hrmgpu_join(self.h))

    def set_scene3d(self, n_arena, n_obs, sq17, n_grid):
        sq17 = _d(sq17).reshape(-1, 17)
        assert sq17.shape[0] == n_arena + n_obs
        self.n_grid = n_grid
        self._chk(self.lib.hrmgpu_set_scene3d(self.h, n_arena, n_obs, _p(sq17), n_grid))

    def set_scene2d(self, n_arena, n_obs, se6, num):
        se6 = _d(se6).reshape(-1, 6)
        assert se6.shape[0] == n_arena + n_obs
        self._chk(self.lib.hrmgpu_set_scene2d(self.h, n_arena, n_obs, _p(se6), num))

    def set_sweep(self, lim, Lx, Ly):
        lim = _d(lim)
        self._chk(self.lib.hrmgpu_set_sweep(self.h, _p(lim), Lx, Ly))

    def resweep(self, lim, Lx, Ly):
        """hrmgpu_resweep: redo sweep + free segments of all built layers on their cached boundaries for a new grid."""
        lim = _d(lim)
        self._chk(self.lib.hrmgpu_resweep(self.h, _p(lim), Lx, Ly))

    # ---- builds ----
    def build_layers(self, body_semi, body_R, body_off, precision=F64_STRICT):
        body_semi = _d(body_semi).reshape(-1, 3)
        B = body_semi.shape[0]
        body_R = _d(body_R).reshape(-1, B, 9)
        K = body_R.shape[0]
        body_off = _d(body_off).reshape(K, B, 3)
        self._chk(self.lib.hrmgpu_build_layers(self.h, K, B, _p(body_semi), _p(body_R), _p(body_off), precision))

    def build_layers_dev(self, K, B, d_semi, d_R, d_off, precision=F64_STRICT):
        self._chk(self.lib.hrmgpu_build_layers_dev(self.h, K, B, C.c_void_p(d_semi), C.c_void_p(d_R), C.c_void_p(d_off), precision))

    def build_layers2d(self, body_semi, body_angle, body_off, precision=F64_STRICT):
        body_semi = _d(body_semi).reshape(-1, 2)
        B = body_semi.shape[0]
        body_angle = _d(body_angle).reshape(-1, B)
        K = body_angle.shape[0]
        body_off = _d(body_off).reshape(K, B, 2)
        self._chk(self.lib.hrmgpu_build_layers2d(self.h, K, B, _p(body_semi), _p(body_angle), _p(body_off), precision))

    # ---- results ----
    def dims(self):
        out = np.zeros(8, dtype=np.int64)
        self._chk(self.lib.hrmgpu_get_dims(self.h, _p(out)))
        return dict(zip(["K", "B", "S", "Sa", "L", "M", "segments", "dim"], (int(v) for v in out)))

    def boundary(self, k, surface):
        d = self.dims()
        out = np.zeros((d["M"], d["dim"]))
        self._chk(self.lib.hrmgpu_get_boundary(self.h, k, surface, _p(out)))
        return out.T.copy()  # dim x M like the reference's BoundaryPoints

    def boundary_layer(self, k):
        """All S boundaries of layer k in one call: [S, dim, M] (each dim x M like the reference's BoundaryPoints)."""
        d = self.dims()
        out = np.zeros((d["S"], d["M"], d["dim"]))
        self._chk(self.lib.hrmgpu_get_boundary_layer(self.h, k, _p(out)))
        return np.ascontiguousarray(out.transpose(0, 2, 1))

    def intervals(self, k):
        d = self.dims()
        lo = np.zeros((d["L"], d["S"]))
        up = np.zeros((d["L"], d["S"]))
        self._chk(self.lib.hrmgpu_get_intervals(self.h, k, _p(lo), _p(up)))
        return lo, up

    def segments(self, k):
        d = self.dims()
        off = np.zeros(d["L"] + 1, dtype=np.int32)
        cnt = self._chk(self.lib.hrmgpu_get_segments(self.h, k, _p(off), None, None, None, 0))
        xL, xU, xM = np.zeros(cnt), np.zeros(cnt), np.zeros(cnt)
        if cnt:
            self._chk(self.lib.hrmgpu_get_segments(self.h, k, _p(off), _p(xL), _p(xU), _p(xM), cnt))
        return off, xL, xU, xM

    def segments_all(self):
        d = self.dims()
        off = np.zeros(d["K"] * d["L"] + 1, dtype=np.int64)
        tot = d["segments"]
        xL, xU, xM = np.zeros(tot), np.zeros(tot), np.zeros(tot)
        self._chk(self.lib.hrmgpu_get_segments_all(self.h, _p(off), _p(xL), _p(xU), _p(xM), tot))
        return off, xL, xU, xM

    def fetch_segments_async(self, off, xL, xU, xM):
        """Queues the read-back of the last build into the given (pinned) numpy arrays and returns; see wait_segments()."""
        assert off.dtype == np.int64 and xL.dtype == xU.dtype == xM.dtype == np.float64
        return self._chk(self.lib.hrmgpu_fetch_segments_async(self.h, _p(off), _p(xL), _p(xU), _p(xM), min(len(xL), len(xU), len(xM))))

    def wait_segments(self):
        """Blocks until the queued read-back is complete; returns the segment count (HrmGpuError E_CAP if it did not fit)."""
        return self._chk(self.lib.hrmgpu_wait_segments(self.h))

    def single_hit_count(self):
        out = np.zeros(1, dtype=np.int64)
        self._chk(self.lib.hrmgpu_get_single_hit_count(self.h, _p(out)))
        return int(out[0])

    def test_edges(self, k, v1, v2):
        d = self.dims()
        v1 = _d(v1).reshape(-1, d["dim"])
        v2 = _d(v2).reshape(-1, d["dim"])
        out = np.zeros(v1.shape[0], dtype=np.uint8)
        if v1.shape[0]:
            self._chk(self.lib.hrmgpu_test_edges(self.h, k, v1.shape[0], _p(v1), _p(v2), _p(out)))
        return out.astype(bool)

    def test_edges_multi(self, layer, v1, v2):
        """Edge e against the C-obstacles of layer layer[e]; one launch for all layers."""
        d = self.dims()
        v1 = _d(v1).reshape(-1, d["dim"])
        v2 = _d(v2).reshape(-1, d["dim"])
        layer = np.ascontiguousarray(layer, dtype=np.int32)
        assert layer.shape[0] == v1.shape[0]
        out = np.zeros(v1.shape[0], dtype=np.uint8)
        if v1.shape[0]:
            self._chk(self.lib.hrmgpu_test_edges_multi(self.h, v1.shape[0], _p(layer), _p(v1), _p(v2), _p(out)))
        return out.astype(bool)

    def build_bridge(self, tfe_semi, tfe_R, precision=F64_STRICT):
        tfe_semi = _d(tfe_semi)
        P, B = tfe_semi.shape[0], tfe_semi.shape[1]
        tfe_R = _d(tfe_R).reshape(P, B, 9)
        self._chk(self.lib.hrmgpu_build_bridge(self.h, P, B, _p(tfe_semi), _p(tfe_R), precision))

    def build_bridge_tfe(self, body_semi, quat_a, quat_b, num_step, precision=F64_STRICT):
        """computeTFE + bridgeSlice on the device (hrmgpu_build_bridge_tfe): body_semi [B][3], quat_a / quat_b [P][B][4];
        returns the fitted ellipsoids [P][B][7] (semi-axes, quaternion)."""
        body_semi = _d(body_semi).reshape(-1, 3)
        B = body_semi.shape[0]
        quat_a, quat_b = _d(quat_a).reshape(-1, B, 4), _d(quat_b).reshape(-1, B, 4)
        P = quat_a.shape[0]
        out = np.zeros((P, B, 7))
        self._chk(self.lib.hrmgpu_build_bridge_tfe(self.h, P, B, _p(body_semi), _p(quat_a), _p(quat_b), int(num_step), precision, _p(out)))
        return out

    def build_bridge2d(self, tfe_semi, tfe_angle, precision=F64_STRICT):
        tfe_semi = _d(tfe_semi)
        P, B = tfe_semi.shape[0], tfe_semi.shape[1]
        tfe_angle = _d(tfe_angle).reshape(P, B)
        self._chk(self.lib.hrmgpu_build_bridge2d(self.h, P, B, _p(tfe_semi), _p(tfe_angle), precision))

    def test_bridge(self, p, centres):
        centres = _d(centres)
        Q = centres.shape[0]
        out = np.zeros(Q, dtype=np.uint8)
        if Q:
            self._chk(self.lib.hrmgpu_test_bridge(self.h, p, Q, _p(centres), _p(out)))
        return out.astype(bool)

    def test_bridge_multi(self, pair, centres):
        """Pose q against the bridge layer of pair pair[q]; one launch for all pairs."""
        centres = _d(centres)
        pair = np.ascontiguousarray(pair, dtype=np.int32)
        Q = centres.shape[0]
        assert pair.shape[0] == Q
        out = np.zeros(Q, dtype=np.uint8)
        if Q:
            self._chk(self.lib.hrmgpu_test_bridge_multi(self.h, Q, _p(pair), _p(centres), _p(out)))
        return out.astype(bool)

    def test_transitions(self, pair, v0, v1, w0, w1, link_off, chain_off=None):
        """isMultiSliceTransitionFree for candidate vertex pairs, interpolation on the device (see include/hrmgpu.h);
        chain_off selects the articulated variant (body centre = link_off + (chain_off + t))."""
        pair = np.ascontiguousarray(pair, dtype=np.int32)
        v0, v1, w0, w1, link_off = _d(v0), _d(v1), _d(w0), _d(w1), _d(link_off)
        Cn = pair.shape[0]
        out = np.zeros(Cn, dtype=np.uint8)
        if Cn and chain_off is None:
            self._chk(self.lib.hrmgpu_test_transitions(self.h, Cn, _p(pair), _p(v0), _p(v1), w0.shape[0], _p(w0), _p(w1), _p(link_off), _p(out)))
        elif Cn:
            chain_off = _d(chain_off)
            self._chk(self.lib.hrmgpu_test_transitions_articulated(self.h, Cn, _p(pair), _p(v0), _p(v1), w0.shape[0], _p(w0), _p(w1), _p(link_off),
                                                                   _p(chain_off), _p(out)))
        return out.astype(bool)

    def connect_slices3d(self, cap):
        """connectOneSlice's pair codes for all layers of the last build (see include/hrmgpu.h): returns (codes, layer_first)."""
        K = self.dims()["K"]
        code = np.zeros(max(1, int(cap)), dtype=np.uint8)
        first = np.zeros(K + 1, dtype=np.int64)
        n = self._chk(self.lib.hrmgpu_connect_slices3d(self.h, _p(code), int(cap), _p(first)))
        return code[:n], first

    def connect_pairs3d(self, vertex_xyz, range4, thx, thy, w0, w1, link_off, chain_off=None):
        """connectMultiSlice's vertex matching on the device (see include/hrmgpu.h): vertex_xyz [V][3], range4 [P][4] =
        {near begin, near end, cur begin, cur end}; returns (match, transition tests made), match[first(p) + (m0 - near begin)] =
        id of the vertex that m0 connects to, or -1."""
        xyz = _d(vertex_xyz).reshape(-1, 3)
        r4 = np.ascontiguousarray(range4, dtype=np.int32).reshape(-1, 4)
        w0, w1, link_off = _d(w0), _d(w1), _d(link_off)
        chain_off = None if chain_off is None else _d(chain_off)
        match = np.full(max(1, int((r4[:, 1] - r4[:, 0]).sum())), -1, dtype=np.int32)
        ncand = C.c_longlong(0)
        n = self._chk(self.lib.hrmgpu_connect_pairs3d(self.h, xyz.shape[0], _p(xyz), r4.shape[0], _p(r4), float(thx), float(thy), w0.shape[0], _p(w0),
                                                      _p(w1), _p(link_off), None if chain_off is None else _p(chain_off), _p(match), match.shape[0],
                                                      C.byref(ncand)))
        return match[:n], int(ncand.value)

    def pack_segments_dev(self, d_ptr=None, cap_bytes=0):
        return self._chk(self.lib.hrmgpu_pack_segments_dev(self.h, C.c_void_p(d_ptr) if d_ptr else None, cap_bytes))

    # ---- multi-GPU exchange ----
    def comm_unique_id(self):
        buf = np.zeros(128, dtype=np.uint8)
        rc = self.lib.hrmgpu_comm_unique_id(_p(buf))
        if rc != OK:
            raise HrmGpuError(rc, "hrmgpu_comm_unique_id failed (NCCL not loadable?)")
        return buf

    def comm_init(self, n_ranks, rank, id128):
        id128 = np.ascontiguousarray(id128, dtype=np.uint8)
        assert id128.size == 128
        self._chk(self.lib.hrmgpu_comm_init(self.h, n_ranks, rank, _p(id128)))

    def comm_destroy(self):
        self._chk(self.lib.hrmgpu_comm_destroy(self.h))

    def allgather_layers(self, line_cap, seg_cap, nccl_comm=None, cuda_stream=None):
        self._chk(self.lib.hrmgpu_allgather_layers(self.h, C.c_void_p(nccl_comm) if nccl_comm else None,
                                                   C.c_void_p(cuda_stream) if cuda_stream else None, line_cap, seg_cap))

    def gathered_info(self):
        out = np.zeros(4, dtype=np.int64)
        ptr = C.c_void_p()
        self._chk(self.lib.hrmgpu_gathered_info(self.h, _p(out), C.byref(ptr)))
        return {"world": int(out[0]), "rank": int(out[1]), "line_cap": int(out[2]), "seg_cap": int(out[3]), "d_blocks": ptr.value}

    def gathered(self, rank):
        """(off[lines+1], xL, xU, xM) of rank `rank` from the last exchange."""
        info = self.gathered_info()
        off = np.zeros(info["line_cap"] + 1, dtype=np.int64)
        lines = C.c_longlong(0)
        tot = self._chk(self.lib.hrmgpu_get_gathered(self.h, rank, C.byref(lines), _p(off), None, None, None, 0))
        xL, xU, xM = np.zeros(tot), np.zeros(tot), np.zeros(tot)
        if tot:
            self._chk(self.lib.hrmgpu_get_gathered(self.h, rank, C.byref(lines), _p(off), _p(xL), _p(xU), _p(xM), tot))
        return off[: lines.value + 1].copy(), xL, xU, xM

    def last_segments_ms(self):
        return float(self.lib.hrmgpu_last_segments_ms(self.h))

    def segment_redos(self, reset=False):
        return int(self.lib.hrmgpu_segment_redos(self.h, int(reset)))

    def kernel_launches(self, reset=False):
        return int(self.lib.hrmgpu_kernel_launches(self.h, int(reset)))

    def last_minksweep_ms(self):
        return float(self.lib.hrmgpu_last_minksweep_ms(self.h))
