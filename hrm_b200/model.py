"""Host-side mirror (Python binding) of the reference's geometric model classes for the hot path:
SuperQuadrics / SuperEllipse parameter holders and MultiBodyTree3D / MultiBodyTree2D, with the same
names and argument meaning as the reference (include/hrm/geometry/SuperQuadrics.h:23-74,
SuperEllipse.h:23-65, include/hrm/datastructure/MultiBodyTree.h:12-63).  They hold parameters and
produce the flat arrays the C ABI consumes; all boundary arithmetic happens on the GPU.
"""
import math

from . import hostmath as hm


class SuperQuadrics:
    def __init__(self, semiAxis, epsilon, position, quat, num):
        self.semiAxis = [float(v) for v in semiAxis]
        self.epsilon = [float(v) for v in epsilon]
        self.position = [float(v) for v in position]
        self.quat = [float(v) for v in quat]  # (w, x, y, z), not normalised (reference keeps it raw)
        self.num = int(num)

    def getNumParam(self):
        return self.num

    def getNum(self):
        return self.num * self.num

    def row17(self):
        return self.semiAxis + self.epsilon + self.position + hm.flat9(hm.quat_to_rot(self.quat))


class SuperEllipse:
    def __init__(self, semiAxis, epsilon, position, angle, num):
        self.semiAxis = [float(v) for v in semiAxis]
        self.epsilon = float(epsilon)
        self.position = [float(v) for v in position]
        self.angle = float(angle)
        self.num = int(num)

    def row6(self):
        return self.semiAxis + [self.epsilon] + self.position + [self.angle]


class MultiBodyTree3D:
    """reference src/datastructure/MultiBodyTree3D.cpp"""

    def __init__(self, base):
        self.base = SuperQuadrics(base.semiAxis, base.epsilon, base.position, base.quat, base.num)
        self.link = []
        self.tf = []  # (R, t) of every link relative to the base

    def addBody(self, link):  # :21-32
        self.link.append(SuperQuadrics(link.semiAxis, link.epsilon, link.position, link.quat, link.num))
        self.tf.append((hm.quat_to_rot(link.quat), list(link.position)))

    def getNumLinks(self):
        return len(self.link)

    def robotTF(self, R, t):  # :34-53, g = [R t; 0 1]
        self.base.position = [float(v) for v in t]
        self.base.quat = hm.rot_to_quat(R)
        for i, (Rl, tl) in enumerate(self.tf):
            Rg = hm.mat_mul(R, Rl)
            tg = [a + b for a, b in zip(hm.mat_vec(R, tl), t)]
            self.link[i].position = tg
            self.link[i].quat = hm.rot_to_quat(Rg)

    def body_arrays(self):
        """(semi [B][3], R [B][9], off [B][3]) of the current pose: the Minkowski sum of body b uses its
        rotation, and link centres are subtracted from the link boundaries (:91-105)."""
        semi = [self.base.semiAxis] + [l.semiAxis for l in self.link]
        R = [hm.flat9(hm.quat_to_rot(self.base.quat))] + [hm.flat9(hm.quat_to_rot(l.quat)) for l in self.link]
        off = [[0.0, 0.0, 0.0]] + [list(l.position) for l in self.link]
        return semi, R, off


class MultiBodyTree2D:
    """reference src/datastructure/MultiBodyTree2D.cpp"""

    def __init__(self, base):
        self.base = SuperEllipse(base.semiAxis, base.epsilon, base.position, base.angle, base.num)
        self.link = []
        self.tf = []

    def addBody(self, link):  # :10-22
        self.link.append(SuperEllipse(link.semiAxis, link.epsilon, link.position, link.angle, link.num))
        c, s = math.cos(link.angle), math.sin(link.angle)
        self.tf.append(([[c, -s], [s, c]], list(link.position)))

    def getNumLinks(self):
        return len(self.link)

    def robotTF(self, theta, t=(0.0, 0.0)):  # :24-41 with g = [Rot(theta) t; 0 1]
        c, s = math.cos(theta), math.sin(theta)
        R = [[c, -s], [s, c]]
        self.base.position = [float(t[0]), float(t[1])]
        self.base.angle = math.atan2(R[1][0], R[0][0])
        for i, (Rl, tl) in enumerate(self.tf):
            m00 = R[0][0] * Rl[0][0] + R[0][1] * Rl[1][0]
            m10 = R[1][0] * Rl[0][0] + R[1][1] * Rl[1][0]
            self.link[i].position = [(R[0][0] * tl[0] + R[0][1] * tl[1]) + t[0], (R[1][0] * tl[0] + R[1][1] * tl[1]) + t[1]]
            self.link[i].angle = math.atan2(m10, m00)

    def body_arrays(self):
        """(semi [B][2], angle [B], off [B][2]) as MultiBodyTree2D::minkSum derives them (:43-65):
        link angle = angle(Rot(base) * R_link), offset = Rot(base) * t_link."""
        semi = [self.base.semiAxis] + [l.semiAxis for l in self.link]
        c, s = math.cos(self.base.angle), math.sin(self.base.angle)
        ang = [self.base.angle]
        off = [[0.0, 0.0]]
        for Rl, tl in self.tf:
            m00 = c * Rl[0][0] + (-s) * Rl[1][0]
            m10 = s * Rl[0][0] + c * Rl[1][0]
            ang.append(math.atan2(m10, m00))
            off.append([c * tl[0] + (-s) * tl[1], s * tl[0] + c * tl[1]])
        return semi, ang, off
