"""Generates tests/golden/yolo26seg_program.json: the statement list + weight literals of the reference's committed
`lele_gen` output (examples/yolo26n-seg/src/yolo26seg.rs), extracted by lele_b200/model_rs.py::parse_model_rs,
plus the non-learned constants a synthetic weights.bin needs (the real blob is not in the checkout).

Run in the build container (reads /root/reference; the GPU box only sees the JSON):
    python tests/golden/make_model_program.py
"""
import importlib.util
import json
import os

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
SRC = "/root/reference/examples/yolo26n-seg/src/yolo26seg.rs"


def main():
    spec = importlib.util.spec_from_file_location("model_rs", os.path.join(ROOT, "lele_b200", "model_rs.py"))   # by path: the package needs the built .so
    m = importlib.util.module_from_spec(spec); spec.loader.exec_module(m)
    prog = m.parse_model_rs(open(SRC).read())
    # constants of the graph that are not learned weights (offsets are the generated file's literals):
    #   attention scale d_k^-1/2 (d_k = 32), Resize scales / sizes, top-k K, class count (index -> (anchor, class) split);
    #   anchor points / strides are rebuilt by the test (grid centres of the 80/40/20 levels, Ultralytics make_anchors)
    prog["constants"] = {"4852416": [32 ** -0.5], "5445328": [1.0, 1.0, 2.0, 2.0], "7762768": [1, 64, 80, 80], "10993152": [300], "10993200": [80]}
    prog["anchor_points_offset"] = 10644560
    prog["anchor_strides_offset"] = 10951184
    prog["source"] = "examples/yolo26n-seg/src/yolo26seg.rs:293-662 (run_chunk_0), weight helpers :676-709"
    out = os.path.join(HERE, "yolo26seg_program.json")
    with open(out, "w") as fh:
        json.dump(prog, fh, separators=(",", ":"))
    print(out, os.path.getsize(out), "bytes;", len(prog["statements"]), "statements")


if __name__ == "__main__":
    main()
