"""Host-side Highway RoadMap planners (Python binding) on top of the C ABI — same names and call structure as the
reference's HighwayRoadMap / HRM3D / HRM2D (include/hrm/planners/HighwayRoadMap-inl.h, src/planners/HRM3D.cpp,
HRM2D.cpp), with the per-layer free-space construction, the intra-layer edge tests and the bridge-layer tests executed
in batches on the GPU:

  buildRoadmap():  ONE hrmgpu_build_layers call for all C-layers (K1+K2+K3), then per layer
      generateVertices  -> from the free-segment CSR
      connectOneSlice   -> every candidate segment of connectOneSlice2D/3D and the four detour segments of bridgeVertex
                           are evaluated speculatively in one hrmgpu_test_edges batch; the reference's loops are then
                           REPLAYED on the host with those booleans, so vertex ids and edge order are the reference's
                           (bridgeVertex appends vertices in the middle of a layer, SURVEY.md finding 8)
  connectMultiSlice(): TFE on the host (tfe.py), ONE hrmgpu_build_bridge call for all layer pairs, all near vertex pairs of
                           a layer pair tested in one hrmgpu_test_bridge batch, loops replayed on the host
  search():        A* on the host (stays on the host per BASELINE.json north_star).

`per_orientation=False` feeds every layer the as-loaded robot pose, which is what the reference snapshot computes
(its FreeSpace robot copy is never re-oriented, SURVEY.md finding 1); `True` gives the intended per-orientation layers.
"""
import heapq
import math
import time

import numpy as np

from . import capi
from . import hostmath as hm
from . import tfe
from .model import MultiBodyTree2D, MultiBodyTree3D


def vector_euclidean(v1, v2):  # src/util/DistanceMetric.cpp:5-13
    acc = 0.0
    for a, b in zip(v1, v2):
        d = a - b
        acc = acc + d * d
    return math.sqrt(acc)


class PlannerParameter:  # include/hrm/planners/PlanningRequest.h:10-32
    def __init__(self):
        self.boundaryLimits = []
        self.numSlice = 0
        self.numLineX = 0
        self.numLineY = 0
        self.numPoint = 5
        self.numSearchNeighbor = 10
        self.searchRadius = hm.HALF_PI


class PlanningRequest:  # :35-47
    def __init__(self):
        self.isRobotRigid = True
        self.parameters = PlannerParameter()
        self.start = []
        self.goal = []


class PlanningResult:  # include/hrm/planners/PlanningResult.h
    def __init__(self):
        self.solved = False
        self.buildTime = self.searchTime = self.totalTime = 0.0
        self.vertex = []   # graphStructure.vertex
        self.edge = []     # graphStructure.edge
        self.weight = []   # graphStructure.weight
        self.PathId = []
        self.solvedPath = []
        self.cost = 0.0


def astar(vertex, edge, weight, s, t):
    """A* with the Euclidean heuristic of HighwayRoadMap-inl.h:135-147; returns the predecessor map or None."""
    n = len(vertex)
    adj = [[] for _ in range(n)]
    for (a, b), w in zip(edge, weight):
        adj[a].append((b, w))
        adj[b].append((a, w))
    dist = [math.inf] * n
    pred = [n] * n
    dist[s] = 0.0
    pred[s] = s
    pq = [(vector_euclidean(vertex[s], vertex[t]), s)]
    closed = [False] * n
    while pq:
        _, u = heapq.heappop(pq)
        if closed[u]:
            continue
        closed[u] = True
        if u == t:
            return pred
        for v, w in adj[u]:
            nd = dist[u] + w
            if nd < dist[v]:
                dist[v] = nd
                pred[v] = u
                heapq.heappush(pq, (nd + vector_euclidean(vertex[v], vertex[t]), v))
    return None


class HighwayRoadMap:
    """Template base of the reference (HighwayRoadMap-inl.h): plan / buildRoadmap / search / bridgeVertex replay."""

    def __init__(self, robot, arena, obstacle, req, device=0, precision=capi.F64_STRICT, per_orientation=False, ctx=None):
        self.robot_ = robot
        self.arena_ = arena
        self.obs_ = obstacle
        self.start_ = list(req.start)
        self.goal_ = list(req.goal)
        self.param_ = req.parameters
        self.res_ = PlanningResult()
        self.vertexIdx_ = []  # per layer: (startId, slice)
        self.precision = precision
        self.per_orientation = per_orientation
        self.ctx = ctx if ctx is not None else capi.Context(device)
        self.numEdgeTests = 0
        self.timing = {}

    def getPlanningResult(self):
        return self.res_

    def getPlannerParameters(self):
        return self.param_

    # :65-89 (no refinement loop: the reference's refinement throws, SURVEY.md finding 6)
    def plan(self, timeLim=None):
        t0 = time.perf_counter()
        self.buildRoadmap()
        self.res_.buildTime = time.perf_counter() - t0
        t0 = time.perf_counter()
        self.search()
        self.res_.searchTime = time.perf_counter() - t0
        self.res_.totalTime = self.res_.buildTime + self.res_.searchTime
        if self.res_.solved:
            pose = len(self.res_.vertex[0])
            self.res_.solvedPath = [self.start_[:pose]] + [self.res_.vertex[i] for i in self.res_.PathId] + [self.goal_[:pose]]

    # :110-175
    def search(self):
        r = self.res_
        if not r.vertex:
            return
        idx_s = self.getNearestNeighborsOnGraph(self.start_, self.param_.numSearchNeighbor, self.param_.searchRadius)
        idx_g = self.getNearestNeighborsOnGraph(self.goal_, self.param_.numSearchNeighbor, self.param_.searchRadius)
        for s in idx_s:
            for g in idx_g:
                pred = astar(r.vertex, r.edge, r.weight, s, g)
                if pred is not None:
                    path, cur, cost = [g], g, 0.0
                    while cur != s:
                        cost += vector_euclidean(r.vertex[cur], r.vertex[pred[cur]])
                        cur = pred[cur]
                        path.append(cur)
                    r.PathId = path[::-1]
                    r.cost = cost
                    r.solved = True
                    return

    # ---- replay helpers ----------------------------------------------------------------------------------------
    def _add_edge(self, i, j):
        self.res_.edge.append((i, j))
        self.res_.weight.append(vector_euclidean(self.res_.vertex[i], self.res_.vertex[j]))

    def _connect_pair(self, idx1, idx2, code, zi):
        """One (idx1, idx2) candidate of connectOneSlice: code 0 = direct edge, 1 / 2 = via the bridge vertex
        vNew1 / vNew2 (bridgeVertex, :295-326), 3 = nothing.  zi = index of the coordinate that is swapped."""
        if code == 0:
            self._add_edge(idx1, idx2)
        elif code in (1, 2):
            v1, v2 = self.res_.vertex[idx1], self.res_.vertex[idx2]
            if code == 1:
                vn = list(v1)
                vn[zi] = v2[zi]
            else:
                vn = list(v2)
                vn[zi] = v1[zi]
            idn = len(self.res_.vertex)
            self.res_.vertex.append(vn)
            self.res_.edge.append((idx1, idn))
            self.res_.weight.append(vector_euclidean(v1, vn))
            self.res_.edge.append((idn, idx2))
            self.res_.weight.append(vector_euclidean(vn, v2))

    @staticmethod
    def _pair_codes(free5):
        """free5: [P][5] booleans of the segments (v1,v2), (v1,n1), (n1,v2), (v1,n2), (n2,v2)."""
        f = np.asarray(free5, dtype=bool)
        code = np.full(len(f), 3, dtype=np.int8)
        code[f[:, 3] & f[:, 4]] = 2
        code[f[:, 1] & f[:, 2]] = 1
        code[f[:, 0]] = 0
        return code


# ==================================================================================================================
class HRM3D(HighwayRoadMap):
    """src/planners/HRM3D.cpp"""

    def __init__(self, robot, arena, obstacle, req, quat_samples, **kw):
        super().__init__(robot, arena, obstacle, req, **kw)
        self.q_ = [list(q) for q in quat_samples]
        self.param_.numSlice = len(self.q_)
        n = arena[0].num
        sq17 = np.array([o.row17() for o in list(arena) + list(obstacle)], dtype=np.float64)
        self.ctx.set_scene3d(len(arena), len(obstacle), sq17, n)
        self.n_grid = n

    # ---- K1+K2+K3 for all layers ----
    def _layer_poses(self):
        semi = Rs = offs = None
        Rs, offs, base_q = [], [], []
        rb = MultiBodyTree3D(self.robot_.base)
        for l in self.robot_.link:
            rb.addBody(l)
        frozen = self.robot_.body_arrays()
        for q in self.q_:
            rb.robotTF(hm.quat_to_rot(q), [0.0, 0.0, 0.0])  # HRM3D.cpp:27-31,510-518 (planner's robot)
            base_q.append(list(rb.base.quat))
            if self.per_orientation:
                semi, R, off = rb.body_arrays()
            else:
                semi, R, off = frozen  # the FreeSpace copy keeps the as-loaded pose (finding 1)
            Rs.append(R)
            offs.append(off)
        return np.array(semi), np.array(Rs), np.array(offs), base_q

    def buildRoadmap(self):
        p = self.param_
        lim, Lx, Ly = p.boundaryLimits, int(p.numLineX), int(p.numLineY)
        t0 = time.perf_counter()
        self.ctx.set_sweep(lim, Lx, Ly)
        semi, R, off, base_q = self._layer_poses()
        self.ctx.build_layers(semi, R, off, self.precision)
        seg_off, xL, xU, xM = self.ctx.segments_all()
        self.timing["layers_s"] = time.perf_counter() - t0
        dx = (lim[1] - lim[0]) / float(Lx - 1)  # HRM3D.cpp:66-81
        dy = (lim[3] - lim[2]) / float(Ly - 1)
        tx = [lim[0] + float(i) * dx for i in range(Lx)]
        ty = [lim[2] + float(i) * dy for i in range(Ly)]
        L = Lx * Ly
        t0 = time.perf_counter()
        for k in range(len(self.q_)):
            o = seg_off[k * L:(k + 1) * L + 1]
            self._connect_one_slice3d(k, base_q[k], tx, ty, o, xL, xU, xM)
        self.timing["connect_one_slice_s"] = time.perf_counter() - t0
        t0 = time.perf_counter()
        self.connectMultiSlice()
        self.timing["connect_multi_slice_s"] = time.perf_counter() - t0

    def _connect_one_slice3d(self, k, q, tx, ty, o, xL, xU, xM):
        """generateVertices + connectOneSlice2D per plane + cross-plane connections (HRM3D.cpp:93-156)."""
        Lx, Ly = len(tx), len(ty)
        qv = [q[0], q[1], q[2], q[3]]

        def seg(ix, iy):
            l = ix * Ly + iy
            return int(o[l]), int(o[l + 1])

        # ---- speculative batch: all candidate pairs of the layer with their 5 segments ----
        pairs = []  # (kind, ix, iy, j1, j2): kind 0 = adjacent lines in plane ix, 1 = adjacent planes at line iy
        P1, P2 = [], []
        for ix in range(Lx):
            for iy in range(Ly - 1):
                a0, a1 = seg(ix, iy)
                b0, b1 = seg(ix, iy + 1)
                for j1 in range(a1 - a0):
                    for j2 in range(b1 - b0):
                        pairs.append((0, ix, iy, j1, j2))
                        P1.append((tx[ix], ty[iy], xM[a0 + j1]))
                        P2.append((tx[ix], ty[iy + 1], xM[b0 + j2]))
        for ix in range(Lx - 1):
            for iy in range(Ly):
                a0, a1 = seg(ix, iy)
                b0, b1 = seg(ix + 1, iy)
                for k1 in range(a1 - a0):
                    for k2 in range(b1 - b0):
                        pairs.append((1, ix, iy, k1, k2))
                        P1.append((tx[ix], ty[iy], xM[a0 + k1]))
                        P2.append((tx[ix + 1], ty[iy], xM[b0 + k2]))
        codes = {}
        if pairs:
            v1 = np.array(P1, dtype=np.float64)
            v2 = np.array(P2, dtype=np.float64)
            n1 = v1.copy()
            n1[:, 2] = v2[:, 2]  # vNew1 = v1 with z of v2 (HighwayRoadMap-inl.h:301-304)
            n2 = v2.copy()
            n2[:, 2] = v1[:, 2]
            A = np.concatenate([v1, v1, n1, v1, n2])
            Bq = np.concatenate([v2, n1, v2, n2, v2])
            free = self.ctx.test_edges(k, A, Bq).reshape(5, len(pairs)).T
            self.numEdgeTests += 5 * len(pairs)
            for pr, c in zip(pairs, self._pair_codes(free)):
                codes[pr] = int(c)

        # ---- replay in the reference's order ----
        V = self.res_.vertex
        start_id = len(V)
        line_ids = []  # numVertex_.line[ix][iy]
        for ix in range(Lx):
            plane = []
            for iy in range(Ly):  # generateVertices
                plane.append(len(V))
                a0, a1 = seg(ix, iy)
                for j in range(a0, a1):
                    V.append([tx[ix], ty[iy], xM[j]] + qv)
            line_ids.append(plane)
            for iy in range(Ly):  # connectOneSlice2D (HighwayRoadMap-inl.h:218-261)
                n1 = plane[iy]
                a0, a1 = seg(ix, iy)
                cnt = a1 - a0
                for j1 in range(cnt):
                    if j1 != cnt - 1 and math.fabs(xU[a0 + j1] - xL[a0 + j1 + 1]) < 1e-6:
                        self._add_edge(n1 + j1, n1 + j1 + 1)
                    if iy != Ly - 1:
                        n2 = plane[iy + 1]
                        b0, b1 = seg(ix, iy + 1)
                        for j2 in range(b1 - b0):
                            self._connect_pair(n1 + j1, n2 + j2, codes[(0, ix, iy, j1, j2)], 2)
        slice_end = len(V)
        for ix in range(Lx - 1):  # HRM3D.cpp:131-155
            for iy in range(Ly):
                n1, n2 = line_ids[ix][iy], line_ids[ix + 1][iy]
                a0, a1 = seg(ix, iy)
                b0, b1 = seg(ix + 1, iy)
                for k1 in range(a1 - a0):
                    for k2 in range(b1 - b0):
                        self._connect_pair(n1 + k1, n2 + k2, codes[(1, ix, iy, k1, k2)], 2)
        # buildRoadmap records numVertex_.slice AFTER constructOneSlice (HighwayRoadMap-inl.h:100-102): that includes the
        # bridge vertices the cross-plane loop appended
        self.vertexIdx_.append((start_id, len(V)))
        del slice_end

    # ---- bridge layers (HRM3D.cpp:158-224, 301-317, 367-425, 521-539) ----
    def connectMultiSlice(self):
        if len(self.vertexIdx_) == 1:
            return
        p = self.param_
        V = self.res_.vertex
        lim = p.boundaryLimits
        rb = MultiBodyTree3D(self.robot_.base)
        for l in self.robot_.link:
            rb.addBody(l)
        B = 1 + len(rb.link)
        semis = [rb.base.semiAxis] + [l.semiAxis for l in rb.link]
        nearest = []
        tfe_semi, tfe_R = [], []
        for i in range(len(self.vertexIdx_)):
            min_dist, min_idx = math.inf, 0
            j = 0
            while j != i and j < p.numSlice:  # only earlier layers (App. C9)
                d = tfe.angular_distance(self.q_[i], self.q_[j])
                if d < min_dist:
                    min_dist, min_idx = d, j
                j += 1
            nearest.append(min_idx)
            ps, pr = [], []
            for b in range(B):  # computeTFE
                if b == 0:
                    qa, qb = self.q_[i], self.q_[min_idx]
                else:
                    Rl = rb.tf[b - 1][0]
                    qa = hm.rot_to_quat(hm.mat_mul(hm.quat_to_rot(self.q_[i]), Rl))
                    qb = hm.rot_to_quat(hm.mat_mul(hm.quat_to_rot(self.q_[min_idx]), Rl))
                s, q = tfe.get_tfe3d(semis[b], qa, qb, int(p.numPoint))
                ps.append(s)
                pr.append(hm.flat9(hm.quat_to_rot(q)))
            tfe_semi.append(ps)
            tfe_R.append(pr)
        self.ctx.build_bridge(np.array(tfe_semi), np.array(tfe_R), self.precision)  # bridgeSlice for every pair
        thx = 2.0 * (lim[1] - lim[0]) / float(p.numLineX)
        thy = 2.0 * (lim[3] - lim[2]) / float(p.numLineY)
        steps = int(p.numPoint)
        for i in range(len(self.vertexIdx_)):
            n22, n_2 = self.vertexIdx_[i]
            start, n2 = self.vertexIdx_[nearest[i]]
            if n2 <= start or n_2 <= n22:
                continue
            A = np.array([V[m][:2] for m in range(start, n2)])
            Bv = np.array([V[m][:2] for m in range(n22, n_2)])
            near = ~((np.abs(A[:, None, 0] - Bv[None, :, 0]) > thx) | (np.abs(A[:, None, 1] - Bv[None, :, 1]) > thy))
            cand = np.argwhere(near)
            free_pair = {}
            if len(cand):
                centres = np.zeros((len(cand), steps + 1, B, 3))
                for ci, (a, b) in enumerate(cand):
                    interp = tfe.interpolate_se3(V[start + a], V[n22 + b], steps)  # numPoint+1 poses incl. both ends
                    for si, vs in enumerate(interp):
                        rb.robotTF(hm.quat_to_rot(vs[3:7]), vs[0:3])  # setTransform
                        centres[ci, si, 0] = rb.base.position
                        for li, l in enumerate(rb.link):
                            centres[ci, si, 1 + li] = l.position
                ok = self.ctx.test_bridge(i, centres.reshape(-1, B, 3)).reshape(len(cand), steps + 1).all(axis=1)
                for (a, b), f in zip(cand, ok):
                    free_pair[(int(a), int(b))] = bool(f)
            for m0 in range(start, n2):  # replay of the reference's loop with its moving lower bound n22
                for m1 in range(n22, n_2):
                    if not near[m0 - start, m1 - self.vertexIdx_[i][0]]:
                        continue
                    if free_pair[(m0 - start, m1 - self.vertexIdx_[i][0])]:
                        self._add_edge(m0, m1)
                        n22 = m1
                        break

    # HRM3D.cpp:443-508 with minQuat initialised to q_[0] (App. C18)
    def getNearestNeighborsOnGraph(self, vertex, k, radius):
        p = self.param_
        V = self.res_.vertex
        qq = vertex[3:7]
        min_quat = self.q_[0]
        min_d = tfe.angular_distance(qq, self.q_[0])
        for q in self.q_:
            d = tfe.angular_distance(qq, q)
            if d < min_d:
                min_d, min_quat = d, q
        quat_list = []
        for q in self.q_:
            if tfe.angular_distance(min_quat, q) < radius:
                quat_list.append(q)
            if len(quat_list) >= k:
                break
        idx = []
        for qc in quat_list:
            idx_slice, best = 0, math.inf
            for i, v in enumerate(V):
                d = vector_euclidean(vertex[:len(v)], v)
                if d < best and tfe.angular_distance(qc, v[3:7]) < 1e-6:
                    best, idx_slice = d, i
            lim = p.boundaryLimits
            if (abs(vertex[0] - V[idx_slice][0]) < radius * (lim[1] - lim[0]) / float(p.numLineX) and
                    abs(vertex[1] - V[idx_slice][1]) < radius * (lim[3] - lim[2]) / float(p.numLineY)):
                idx.append(idx_slice)
        return idx


# ==================================================================================================================
class HRM2D(HighwayRoadMap):
    """src/planners/HRM2D.cpp"""

    def __init__(self, robot, arena, obstacle, req, **kw):
        super().__init__(robot, arena, obstacle, req, **kw)
        se6 = np.array([o.row6() for o in list(arena) + list(obstacle)], dtype=np.float64)
        self.ctx.set_scene2d(len(arena), len(obstacle), se6, arena[0].num)
        self.headings_ = []

    def sampleOrientations(self):  # :46-51
        dr = 2 * hm.PI / (float(self.param_.numSlice) - 1)
        self.headings_ = [-hm.PI + dr * float(i) for i in range(int(self.param_.numSlice))]

    def _robot(self):
        rb = MultiBodyTree2D(self.robot_.base)
        for l in self.robot_.link:
            rb.addBody(l)
        return rb

    def buildRoadmap(self):
        p = self.param_
        self.sampleOrientations()
        lim, Ly = p.boundaryLimits, int(p.numLineY)
        self.ctx.set_sweep(lim, 1, Ly)
        rb = self._robot()
        frozen = self.robot_.body_arrays()
        angs, offs, base_angle = [], [], []
        semi = None
        for th in self.headings_:
            rb.robotTF(th)  # setTransform({0,0,theta}) on the planner's robot
            base_angle.append(rb.base.angle)
            semi, a, o = rb.body_arrays() if self.per_orientation else frozen
            angs.append(a)
            offs.append(o)
        self.ctx.build_layers2d(np.array(semi), np.array(angs), np.array(offs), self.precision)
        seg_off, xL, xU, xM = self.ctx.segments_all()
        dy = (lim[3] - lim[2]) / (float(Ly) - 1)  # HRM2D.cpp:56-61
        ty = [lim[2] + float(i) * dy for i in range(Ly)]
        V = self.res_.vertex
        for k in range(len(self.headings_)):
            o = seg_off[k * Ly:(k + 1) * Ly + 1]
            pairs, P1, P2 = [], [], []
            for iy in range(Ly - 1):
                a0, a1, b0, b1 = int(o[iy]), int(o[iy + 1]), int(o[iy + 1]), int(o[iy + 2])
                for j1 in range(a1 - a0):
                    for j2 in range(b1 - b0):
                        pairs.append((iy, j1, j2))
                        P1.append((xM[a0 + j1], ty[iy]))
                        P2.append((xM[b0 + j2], ty[iy + 1]))
            free = {}
            if pairs:
                # bridgeVertex swaps index 2 = theta in 2D, so vNew1 == v1 and vNew2 == v2 as points: it can only re-test
                # the same segment plus a zero-length one (App. C8); a blocked pair stays unconnected.
                f = self.ctx.test_edges(k, np.array(P1), np.array(P2))
                self.numEdgeTests += len(pairs)
                free = dict(zip(pairs, f.tolist()))
            start_id = len(V)
            plane = []
            for iy in range(Ly):  # generateVertices (:72-89): (xM, ty, theta)
                plane.append(len(V))
                for j in range(int(o[iy]), int(o[iy + 1])):
                    V.append([xM[j], ty[iy], base_angle[k]])
            for iy in range(Ly):  # connectOneSlice2D
                n1 = plane[iy]
                a0, a1 = int(o[iy]), int(o[iy + 1])
                for j1 in range(a1 - a0):
                    if j1 != a1 - a0 - 1 and math.fabs(xU[a0 + j1] - xL[a0 + j1 + 1]) < 1e-6:
                        self._add_edge(n1 + j1, n1 + j1 + 1)
                    if iy != Ly - 1:
                        n2 = plane[iy + 1]
                        for j2 in range(int(o[iy + 2]) - int(o[iy + 1])):
                            if free[(iy, j1, j2)]:
                                self._add_edge(n1 + j1, n2 + j2)
            self.vertexIdx_.append((start_id, len(V)))
        self.connectMultiSlice()

    def connectMultiSlice(self):  # :91-156
        if len(self.vertexIdx_) == 1:
            return
        p = self.param_
        V = self.res_.vertex
        lim = p.boundaryLimits
        rb = self._robot()
        B = 1 + len(rb.link)
        semis = [rb.base.semiAxis] + [l.semiAxis for l in rb.link]
        ns = int(p.numSlice)
        adj, tfe_semi, tfe_ang = [], [], []
        for i in range(len(self.vertexIdx_)):
            j = 0 if (i == ns - 1 and ns != 2) else i + 1
            adj.append(j)
            tha, thb = self.headings_[i], self.headings_[j]
            ps, pa = [], []
            for b in range(B):  # computeTFE (:352-373)
                if b == 0:
                    a_, b_ = tha, thb
                else:
                    Rl = rb.tf[b - 1][0]
                    rl = math.atan2(Rl[1][0], Rl[0][0])
                    cl, sl = math.cos(rl), math.sin(rl)

                    def ang(th):
                        c, s = math.cos(th), math.sin(th)
                        return math.atan2(s * cl + c * sl, c * cl + (-s) * sl)

                    a_, b_ = ang(tha), ang(thb)
                s, an = tfe.get_tfe2d(semis[b], a_, b_, int(p.numPoint))
                ps.append(s)
                pa.append(an)
            tfe_semi.append(ps)
            tfe_ang.append(pa)
        self.ctx.build_bridge2d(np.array(tfe_semi), np.array(tfe_ang), self.precision)
        th = 2.0 * math.fabs(lim[3] - lim[2]) / float(p.numLineY)
        steps = int(p.numPoint)
        dt = 1.0 / (float(steps) - 1)
        for i in range(len(self.vertexIdx_)):
            s_cur, e_cur = self.vertexIdx_[i]
            s_adj0, e_adj = self.vertexIdx_[adj[i]]
            s_adj = s_adj0
            if e_cur <= s_cur or e_adj <= s_adj:
                continue
            A = np.array([V[m][1] for m in range(s_cur, e_cur)])
            Bv = np.array([V[m][1] for m in range(s_adj, e_adj)])
            near = ~(np.abs(A[:, None] - Bv[None, :]) > th)
            cand = np.argwhere(near)
            free_pair = {}
            if len(cand):
                centres = np.zeros((len(cand), steps, B, 2))
                for ci, (a, b) in enumerate(cand):
                    v1, v2 = V[s_cur + a], V[s_adj0 + b]
                    for si in range(steps):  # :237-247
                        vs = [(1.0 - float(si) * dt) * v1[j] + float(si) * dt * v2[j] for j in range(3)]
                        rb.robotTF(vs[2], (vs[0], vs[1]))
                        centres[ci, si, 0] = rb.base.position
                        for li, l in enumerate(rb.link):
                            centres[ci, si, 1 + li] = l.position
                ok = self.ctx.test_bridge(i, centres.reshape(-1, B, 2)).reshape(len(cand), steps).all(axis=1)
                for (a, b), f in zip(cand, ok):
                    free_pair[(int(a), int(b))] = bool(f)
            for m0 in range(s_cur, e_cur):
                for m1 in range(s_adj, e_adj):
                    if not near[m0 - s_cur, m1 - s_adj0]:
                        continue
                    if free_pair[(m0 - s_cur, m1 - s_adj0)]:
                        self._add_edge(m0, m1)
                        s_adj = m1
                        break

    def getNearestNeighborsOnGraph(self, vertex, k, radius):  # :284-342
        p = self.param_
        V = self.res_.vertex
        min_angle = V[0][2]
        min_d = math.fabs(vertex[2] - min_angle)
        for v in V:
            d = math.fabs(vertex[2] - v[2])
            if d < min_d:
                min_d, min_angle = d, v[2]
        ang_list = []
        for h in self.headings_:
            if math.fabs(min_angle - h) < radius:
                ang_list.append(h)
            if len(ang_list) >= k:
                break
        idx = []
        lim = p.boundaryLimits
        for ac in ang_list:
            idx_slice, best = 0, vector_euclidean(vertex, V[0])
            for i, v in enumerate(V):
                d = vector_euclidean(vertex, v)
                if d < best and math.fabs(v[2] - ac) < hm.EPSILON:
                    best, idx_slice = d, i
            if abs(vertex[1] - V[idx_slice][1]) < radius * (lim[1] - lim[0]) / float(p.numLineY):
                idx.append(idx_slice)
        return idx
