"""Extract the NUMERIC model data the hot path needs from the reference's assets into small JSON files.

    python tools/extract_reference_assets.py        (build container only: reads /root/reference)

  xml_models/manipulators/sequential.xml:13-39,80-82  -> rigid_body_manipulation_b200/assets/sequential.json
        (per body: pos / euler, joint type / axis / pos, inertial pos / mass / diaginertia; the attachment site; keyframe)
  xml_models/targets/*/object_cad_gt.csv row 1        -> rigid_body_manipulation_b200/assets/targets.json
        (aabb_scale, total_mass, CoM, inertia about the CoM, principal moments, static-XYZ euler)

Only numbers are taken (no meshes, no code); values are written with repr() so they round-trip bit-exactly.  The
GPU box has no /root/reference, which is why these travel with the repo.
"""
import csv
import json
import os
import xml.etree.ElementTree as ET

REF = os.environ.get("RBM_REFERENCE_ROOT", "/root/reference")
OUT = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "rigid_body_manipulation_b200", "assets")


def floats(s, default):
    return [float(x) for x in s.split()] if s is not None else list(default)


def extract_manipulator(path):
    root = ET.parse(path).getroot()
    links, sites = [], []

    def walk(elem, parent):
        for b in elem.findall("body"):
            j = b.find("joint")
            inert = b.find("inertial")
            rec = {
                "name": b.get("name"), "parent": parent,
                "pos": floats(b.get("pos"), [0, 0, 0]), "euler_deg": floats(b.get("euler"), [0, 0, 0]),
                "joint": {"name": j.get("name"), "type": j.get("type", "hinge"), "axis": floats(j.get("axis"), [0, 0, 1]), "pos": floats(j.get("pos"), [0, 0, 0])},
                "inertial": {"pos": floats(inert.get("pos"), [0, 0, 0]), "mass": float(inert.get("mass")), "diaginertia": floats(inert.get("diaginertia"), [])},
            }
            links.append(rec)
            for s in b.findall("site"):
                sites.append({"name": s.get("name"), "body": b.get("name"), "pos": floats(s.get("pos"), [0, 0, 0]), "euler_deg": floats(s.get("euler"), [0, 0, 0])})
            walk(b, b.get("name"))

    walk(root.find("worldbody"), "world")
    key = root.find("keyframe/key")
    return {
        "source": "xml_models/manipulators/sequential.xml", "angle": "degree", "eulerseq": "xyz", "links": links, "sites": sites,
        "keyframe": {"name": key.get("name"), "qpos": floats(key.get("qpos"), [])},
        "gravity": [0.0, 0.0, -9.81], "timestep": 0.002,
        "ft_sensor_site": {"parent_site": "attachment", "euler_deg": [0.0, 0.0, 180.0], "source": "core/core.py:241-242"},
    }


def extract_targets(tdir):
    out = {}
    for name in sorted(os.listdir(tdir)):
        p = os.path.join(tdir, name, "object_cad_gt.csv")
        if not os.path.isfile(p):
            continue
        with open(p, newline="") as f:
            rd = csv.reader(f)
            header, row = next(rd), next(rd)
        rec = {k: (v if k == "id" else float(v)) for k, v in zip(header, row)}
        out[name] = rec
    return out


def main():
    os.makedirs(OUT, exist_ok=True)
    man = extract_manipulator(os.path.join(REF, "xml_models", "manipulators", "sequential.xml"))
    with open(os.path.join(OUT, "sequential.json"), "w") as f:
        json.dump(man, f, indent=1)
    tg = extract_targets(os.path.join(REF, "xml_models", "targets"))
    with open(os.path.join(OUT, "targets.json"), "w") as f:
        json.dump(tg, f, indent=1)
    print(len(man["links"]), "links;", len(tg), "targets:", ", ".join(tg))


if __name__ == "__main__":
    main()
