"""Generates tests/golden/scenes.json from the reference's bundled resources (run in the build container,
where /root/reference exists; the GPU box only sees the committed JSON).

It applies the reference's own two-stage input pipeline so the numbers are exactly what a reference
planner would hold in memory:
  resources/*.csv --parse2DCsvFile (stof)--> doubles --parseGeometricModel (axis-angle -> quaternion,
  ostream << double = "%g", 6 significant digits)--> config/*.csv --parse2DCsvFile (stof)--> doubles
(reference src/util/Parse2dCsvFile.cpp:3-43, src/test/util/ParsePlanningSettings.cpp:162-263), plus
defineParameters (ParsePlanningSettings.cpp:108-160) for the boundary limits and sweep-line counts.

    python tests/golden/make_scenes.py
"""
import ctypes
import json
import math
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from oracle import hrmo  # noqa: E402  (angle-axis -> quaternion with the oracle's Eigen-exact formula)

REF = "/root/reference/resources"
_libc = ctypes.CDLL(None)
_libc.strtof.restype = ctypes.c_float
_libc.strtof.argtypes = [ctypes.c_char_p, ctypes.c_void_p]


def stof(s):
    return float(_libc.strtof(s.strip().encode(), None))


def parse_csv(path):
    rows = []
    with open(path) as f:
        for line in f.read().split("\n"):
            if line == "" or line[0] == "#":
                if line == "":
                    continue
                continue
            rec = []
            for cell in line.split(","):
                try:
                    if cell.strip() == "":
                        raise ValueError
                    rec.append(stof(cell))
                except ValueError:
                    pass
            rows.append(rec)
    return rows


def g6(x):
    return "%g" % x


def roundtrip(x):
    return stof(g6(x))


def geometric_model(path):
    """parseGeometricModel + re-read."""
    out = []
    for obj in parse_csv(path):
        if len(obj) == 12:
            q = hrmo.angle_axis_to_quat(obj[11], obj[8:11])  # normalises the axis like the reference
            row = [roundtrip(v) for v in obj[:8]] + [roundtrip(v) for v in q]
        else:
            row = [roundtrip(v) for v in obj]
        out.append(row)
    return out


def end_points(path, dim, robot_type):
    out = []
    for cfg in parse_csv(path):
        if dim == "2D":
            out.append([roundtrip(v) for v in cfg[:3]])
        else:
            q = hrmo.angle_axis_to_quat(cfg[6], cfg[3:6])
            cfg = list(cfg)
            cfg[3:7] = list(q)
            ndof = {"snake": 9, "tree": 12}.get(robot_type, 6)
            out.append([roundtrip(v) for v in cfg[: ndof + 1]])
    return out


def scene3d(obj_type, env, robot, quat_file="q_icosahedron_60"):
    arena = geometric_model(f"{REF}/3D/env_{obj_type}_{env}_3D_arena.csv")
    obstacle = geometric_model(f"{REF}/3D/env_{obj_type}_{env}_3D_obstacle.csv")
    rob = geometric_model(f"{REF}/3D/robot_{robot}_3D.csv")
    ends = end_points(f"{REF}/3D/setting_{obj_type}_{env}_3D.csv", "3D", robot)
    quats = [[r[3], r[0], r[1], r[2]] for r in parse_csv(f"{REF}/SO3_sequence/{quat_file}.csv")]  # stored w,x,y,z
    f = 1.2
    bound = [arena[0][i] - f * rob[0][0] for i in range(3)]
    lim = [arena[0][5] - bound[0], arena[0][5] + bound[0], arena[0][6] - bound[1], arena[0][6] + bound[1],
           arena[0][7] - bound[2], arena[0][7] + bound[2]]
    min_obs = min(min(o[:3]) for o in obstacle)
    return {"dim": 3, "arena": arena, "obstacle": obstacle, "robot": rob, "end_points": ends, "quats": quats,
            "lim": lim, "Lx": int(math.floor(bound[0] / min_obs)), "Ly": int(math.floor(bound[1] / min_obs)),
            "source": f"{obj_type}/{env}/{robot}/{quat_file}"}


def scene2d(obj_type, env, robot):
    arena = geometric_model(f"{REF}/2D/env_{obj_type}_{env}_2D_arena.csv")
    obstacle = geometric_model(f"{REF}/2D/env_{obj_type}_{env}_2D_obstacle.csv")
    rob = geometric_model(f"{REF}/2D/robot_{robot}_2D.csv")
    ends = end_points(f"{REF}/2D/setting_{obj_type}_{env}_2D.csv", "2D", robot)
    f = 1.5
    bound = [arena[0][0] - f * rob[0][0], arena[0][1] - f * rob[0][0]]
    lim = [arena[0][3] - bound[0], arena[0][3] + bound[0], arena[0][4] - bound[1], arena[0][4] + bound[1]]
    min_obs = min(min(o[:2]) for o in obstacle)
    return {"dim": 2, "arena": arena, "obstacle": obstacle, "robot": rob, "end_points": ends, "lim": lim,
            "Ly": int(bound[1] / min_obs), "source": f"{obj_type}/{env}/{robot}"}


def emit_urdf(name):
    """Re-emits the kinematic skeleton (joints only: parent, child, origin, axis, type) of resources/3D/urdf/<name>.urdf."""
    import xml.etree.ElementTree as ET
    root = ET.parse(f"{REF}/3D/urdf/{name}.urdf").getroot()
    lines = ['<?xml version="1.0" ?>', f'<!-- kinematic skeleton of the reference robot "{name}" (joints only), written by tests/golden/make_scenes.py -->',
             f'<robot name="{name}">']
    links = []
    for j in root.findall("joint"):
        for tag in ("parent", "child"):
            if j.find(tag).get("link") not in links:
                links.append(j.find(tag).get("link"))
    for link in links:
        lines.append(f'  <link name="{link}"/>')
    for j in root.findall("joint"):
        o, a = j.find("origin"), j.find("axis")
        lines.append(f'  <joint name="{j.get("name")}" type="{j.get("type")}">')
        lines.append(f'    <parent link="{j.find("parent").get("link")}"/>')
        lines.append(f'    <child link="{j.find("child").get("link")}"/>')
        lines.append(f'    <origin rpy="{o.get("rpy", "0 0 0")}" xyz="{o.get("xyz", "0 0 0")}"/>')
        if a is not None:
            lines.append(f'    <axis xyz="{a.get("xyz")}"/>')
        lines.append("  </joint>")
    lines.append("</robot>")
    out = os.path.join(os.path.dirname(os.path.abspath(__file__)), f"{name}.urdf")
    with open(out, "w") as f:
        f.write("\n".join(lines) + "\n")
    print("wrote", out)


def main():
    for name in ("snake", "tree"):
        emit_urdf(name)

    scenes = {}
    for env in ["sparse", "cluttered", "maze", "narrow", "home"]:
        scenes[f"3D_superquadrics_{env}_rabbit"] = scene3d("superquadrics", env, "rabbit")
    scenes["3D_superquadrics_home_snake"] = scene3d("superquadrics", "home", "snake")
    scenes["3D_superquadrics_sparse_snake"] = scene3d("superquadrics", "sparse", "snake")
    scenes["3D_superquadrics_sparse_tree"] = scene3d("superquadrics", "sparse", "tree")
    scenes["3D_superquadrics_cluttered_snake"] = scene3d("superquadrics", "cluttered", "snake")
    scenes["3D_ellipsoid_sparse_rabbit"] = scene3d("ellipsoid", "sparse", "rabbit")
    for env in ["sparse", "cluttered", "maze1", "maze2", "maze3", "NAO_lab"]:
        scenes[f"2D_superellipse_{env}_rabbit"] = scene2d("superellipse", env, "rabbit")
    scenes["2D_ellipse_sparse_rabbit"] = scene2d("ellipse", "sparse", "rabbit")
    out = os.path.join(os.path.dirname(os.path.abspath(__file__)), "scenes.json")
    with open(out, "w") as f:
        json.dump(scenes, f, indent=0, sort_keys=True)
    print("wrote", out, {k: (len(v["arena"]), len(v["obstacle"]), len(v["robot"])) for k, v in scenes.items()})


if __name__ == "__main__":
    main()
