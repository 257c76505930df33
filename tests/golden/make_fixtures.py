"""Generate the committed fixtures under tests/golden/ from the read-only reference tree.

Run once in the build container (the GPU box has no /root/reference):

    python tests/golden/make_fixtures.py

What it extracts (data only, no reference source code):
  * the example-model input tables the reference's example generator reads
    (/root/reference/examples/data/input_data/jan_models/model{1,2,5,7}_*.csv,
    cited from gempy/API/examples_generator.py:132-293)  ->  gempy_b200/data/example_inputs.json
  * the four approved scalar-field vectors of
    test/test_model_types/test_example_models_I.py:19-88            ->  approved_scalar_fields.json
  * the Greenstone model the reference ships as examples/data/gempy_models/Greenstone.gempy
    (loaded by gempy/API/examples_generator.py:489-508 through gempy/modules/serialization/save_load.py:168-189):
    input tables + transform + grid  ->  gempy_b200/data/greenstone.json, and the ENGINE OUTPUTS stored in its header
    (``scalar_field_at_interface`` of every element, full double precision)  ->  greenstone_isovalues.json
  * the known answer of test/test_modules/test_grids/test_custom_grid.py:44-47 is a literal
    ([3,3,3,3,1,1,1,1]) and lives in the test itself.
"""
import csv
import json
import os
import re

REF = "/root/reference"
HERE = os.path.dirname(os.path.abspath(__file__))


def read_csv(path):
    with open(path, newline="") as fh:
        rows = list(csv.DictReader(fh))
    return rows


def tables(model):
    base = f"{REF}/examples/data/input_data/jan_models/{model}"
    sp = read_csv(base + "_surface_points.csv")
    ori = read_csv(base + "_orientations.csv")
    out = {
        "surface_points": {
            "X": [float(r["X"]) for r in sp],
            "Y": [float(r["Y"]) for r in sp],
            "Z": [float(r["Z"]) for r in sp],
            "formation": [r["formation"] for r in sp],
        },
        "orientations": {
            "X": [float(r["X"]) for r in ori],
            "Y": [float(r["Y"]) for r in ori],
            "Z": [float(r["Z"]) for r in ori],
            "azimuth": [float(r["azimuth"]) for r in ori],
            "dip": [float(r["dip"]) for r in ori],
            "polarity": [float(r["polarity"]) for r in ori],
            "formation": [r["formation"] for r in ori],
        },
    }
    return out


def approved(name):
    d = f"{REF}/test/test_model_types"
    fn = [f for f in os.listdir(d) if f.endswith(".approved.txt") and name in f][0]
    txt = open(os.path.join(d, fn)).read()
    return [float(x) for x in re.findall(r"[-+]?\d+\.\d+(?:e[-+]?\d+)?", txt)]


def greenstone():
    import zipfile
    import numpy as np
    z = zipfile.ZipFile(f"{REF}/examples/data/gempy_models/Greenstone.gempy")
    h = json.loads(z.read("header.json"))
    raw = z.read("input.bin")
    # table layouts: gempy/core/data/surface_points.py:25, orientations.py:24
    sp_dt = np.dtype([("X", "f8"), ("Y", "f8"), ("Z", "f8"), ("id", "i4"), ("nugget", "f8")])
    ori_dt = np.dtype([("X", "f8"), ("Y", "f8"), ("Z", "f8"), ("G_x", "f8"), ("G_y", "f8"), ("G_z", "f8"), ("id", "i4"), ("nugget", "f8")])
    nb = h["structural_frame"]["binary_meta_data"]
    sp = np.frombuffer(raw[:nb["sp_binary_length"]], dtype=sp_dt)
    ori = np.frombuffer(raw[nb["sp_binary_length"]:nb["sp_binary_length"] + nb["ori_binary_length"]], dtype=ori_dt)
    groups, gold = [], {}
    for g in h["structural_frame"]["structural_groups"]:
        els = []
        for e in g["elements"]:
            s_, o_ = sp[sp["id"] == e["_id"]], ori[ori["id"] == e["_id"]]
            els.append({"name": e["name"],
                        "sp_xyz": np.stack([s_["X"], s_["Y"], s_["Z"]], 1).tolist(), "sp_nugget": s_["nugget"].tolist(),
                        "ori_xyz": np.stack([o_["X"], o_["Y"], o_["Z"]], 1).tolist(),
                        "ori_grad": np.stack([o_["G_x"], o_["G_y"], o_["G_z"]], 1).tolist(), "ori_nugget": o_["nugget"].tolist()})
            gold[e["name"]] = e["scalar_field_at_interface"]
        groups.append({"name": g["name"], "structural_relation": g["structural_relation"], "elements": els})
    model = {"groups": groups, "input_transform": {k: h["input_transform"][k] for k in ("position", "rotation", "scale")},
             "extent": h["grid"]["_octree_grid"]["extent"], "resolution": h["grid"]["_octree_grid"]["resolution"],
             "kernel_options": h["_interpolation_options"]["kernel_options"],
             "number_octree_levels": h["_interpolation_options"]["evaluation_options"]["_number_octree_levels"]}
    return model, gold


def main():
    gs_model, gs_gold = greenstone()
    with open(os.path.join(HERE, "..", "..", "gempy_b200", "data", "greenstone.json"), "w") as fh:
        json.dump(gs_model, fh)
    with open(os.path.join(HERE, "greenstone_isovalues.json"), "w") as fh:
        json.dump(gs_gold, fh, indent=0)
    inputs = {m: tables(m) for m in ("model1", "model2", "model5", "model7")}
    # the example input tables are package data of the host-side example builders
    pkg = os.path.join(HERE, "..", "..", "gempy_b200", "data", "example_inputs.json")
    with open(pkg, "w") as fh:
        json.dump(inputs, fh, indent=0)
    gold = {
        "anticline": approved("Anticline"),
        "fault": approved("Fault Scalar"),
        "combination": approved("Combination"),
        "horizontal_stale": approved("Horizontal"),
    }
    for k, v in gold.items():
        print(k, len(v))
    with open(os.path.join(HERE, "approved_scalar_fields.json"), "w") as fh:
        json.dump(gold, fh, indent=0)


if __name__ == "__main__":
    main()
