"""Golden vectors for the output writers, made by the REFERENCE's own functions: ``nearest_interp`` of
postprocess/postprocess.py and ``au_to_str`` of test_aff2.py are pulled out of the reference sources by name (the files themselves
cannot be imported: both run their script body at import time) and executed on seeded cases.  Run in the build container only:

    python tests/golden/make_golden_outputs.py        # writes tests/golden/outputs.json
"""
import ast
import io
import json
import os
import contextlib

import numpy as np

REF = "/root/reference"
HERE = os.path.dirname(os.path.abspath(__file__))


def function_from(path, name):
    tree = ast.parse(open(path, encoding="utf-8").read())
    for node in tree.body:
        if isinstance(node, ast.FunctionDef) and node.name == name:
            ns = {"np": np}
            exec(compile(ast.Module(body=[node], type_ignores=[]), path, "exec"), ns)
            return ns[name]
    raise KeyError(name)


def main():
    nearest_interp = function_from(os.path.join(REF, "postprocess", "postprocess.py"), "nearest_interp")
    au_to_str = function_from(os.path.join(REF, "test_aff2.py"), "au_to_str")
    rng = np.random.default_rng(20240)
    interp = []
    fixed = [([1, 2, 4, 5], 5), ([1, 2, 3], 3), ([1, 2, 3], 2), ([3, 7, 8, 20], 30), ([5], 9), ([2, 3, 10], 6), ([1, 4, 2, 9], 12)]
    for _ in range(40):
        n_frame = int(rng.integers(1, 80))
        k = int(rng.integers(1, max(2, n_frame)))
        frames = sorted(int(v) for v in rng.choice(np.arange(1, n_frame + 6), size=min(k, n_frame + 5), replace=False))
        fixed.append((frames, n_frame))
    for frames, n_frame in fixed:
        with contextlib.redirect_stdout(io.StringIO()):
            out = nearest_interp(list(frames), n_frame)
        interp.append({"frames": list(frames), "n_frame": n_frame, "indices": [int(v) for v in out]})
    rows = []
    for _ in range(16):
        arr = rng.integers(0, 2, size=12)
        rows.append({"decisions": [int(v) for v in arr], "row": au_to_str(arr)})
    with open(os.path.join(HERE, "outputs.json"), "w") as f:
        json.dump({"nearest_interp": interp, "au_to_str": rows}, f)
    print(len(interp), "interp cases,", len(rows), "rows")


if __name__ == "__main__":
    main()
