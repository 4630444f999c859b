"""Output writers and post-processing against vectors produced by the reference's own functions
(tests/golden/make_golden_outputs.py) and against the file layout of test_aff2.py:80-119.  Host-side code: runs without a GPU."""
import json
import os

import numpy as np
import torch

import avformer_b200 as A
from avformer_b200 import outputs as OUT

GOLD = json.load(open(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "outputs.json")))


def test_nearest_interp_matches_the_reference_function():
    assert len(GOLD["nearest_interp"]) >= 40
    for case in GOLD["nearest_interp"]:
        assert OUT.nearest_interp(case["frames"], case["n_frame"]) == case["indices"], case
    # unsorted input is sorted first, like the reference
    assert OUT.nearest_interp([9, 1, 4, 2], 12) == OUT.nearest_interp([1, 2, 4, 9], 12)


def test_au_rows_match_the_reference_format():
    for case in GOLD["au_to_str"]:
        assert OUT.au_row(case["decisions"]) == case["row"]
    logits = torch.tensor([[0.0, 1e-6, -1e-6, 3.0, -3.0, 0.5, -0.5, 10.0, -10.0, 0.0, 2.0, -2.0] + [7.0] * 9])
    ref = np.round(torch.sigmoid(logits[:, :12]).numpy()).astype(np.int64)          # train.py:155 / test_aff2.py:112-113
    assert (OUT.au_decisions(logits) == ref).all() and ref[0, 0] == 0                # sigmoid(0) = 0.5 rounds to even


def test_result_writer_files_and_pickle(tmp_path):
    g = torch.Generator().manual_seed(5)
    n = 11
    out = torch.zeros(n, 21)
    out[:, :12] = torch.randn(n, 12, generator=g)
    videos = ["v1"] * 4 + ["v2_left"] * 3 + ["v3"] * 4
    order = [3, 0, 1, 2, 6, 5, 4, 10, 9, 8, 7]                                       # dataset indices in sampler order
    with A.AUResultWriter(str(tmp_path), n) as w:
        w.add(videos[:5], order[:5], out[:5])
        w.add(videos[5:], order[5:], out[5:], decisions=torch.from_numpy(OUT.au_decisions(out[5:])).int())
    pk = torch.load(os.path.join(tmp_path, "inference.pkl"))
    assert set(pk) == {"predictions"} and pk["predictions"].shape == (n, 21)
    assert torch.equal(pk["predictions"][torch.tensor(order)], out)
    dec = np.round(torch.sigmoid(out[:, :12]).numpy()).astype(int)
    row = 0
    for vid, cnt in (("v1", 4), ("v2_left", 3), ("v3", 4)):
        lines = open(os.path.join(tmp_path, "au", vid + ".txt")).read().split("\n")
        assert lines[0] == "AU1,AU2,AU4,AU6,AU7,AU10,AU12,AU15,AU23,AU24,AU25,AU26" and lines[-1] == "" and len(lines) == cnt + 2
        for k in range(cnt):
            assert lines[1 + k] == ",".join(str(v) for v in dec[row + k])
        row += cnt


def test_postprocess_expands_to_the_video_length(tmp_path):
    pred_dir, out_dir = tmp_path / "AU", tmp_path / "new"
    os.makedirs(pred_dir)
    frames = {"12-24-1920x1080_left": [2, 3, 7, 8], "clipA": [1, 2, 3]}
    n_frames = {"12-24-1920x1080": 10, "clipA": 3}
    for name, fr in frames.items():
        with open(pred_dir / (name + ".txt"), "w") as f:
            f.write(OUT.HEADERS["AU"] + "\n")
            for k in range(len(fr)):
                f.write(OUT.au_row([(k >> b) & 1 for b in range(12)]) + "\n")
    written = OUT.postprocess_directory(str(pred_dir), str(out_dir), lambda a: frames[a], lambda v: n_frames[v])
    assert written == {"12-24-1920x1080_left.txt": 10, "clipA.txt": 3}
    lines = open(out_dir / "12-24-1920x1080_left.txt").read().split("\n")[:-1]
    which = OUT.nearest_interp(frames["12-24-1920x1080_left"], 10)
    src = open(pred_dir / "12-24-1920x1080_left.txt").read().split("\n")[:-1]
    assert lines[0] == src[0] and lines[1:] == [src[1 + which[i]] for i in range(10)]
    assert which[:10] == [0, 1, 1, 1, 1, 2, 3, 3, 3, 3]
    assert open(out_dir / "clipA.txt").read() == open(pred_dir / "clipA.txt").read()
