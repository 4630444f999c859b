"""Output formats of the reference's inference script and post-processing, fed from the device decisions.

* ``AUResultWriter`` — what ``test_aff2.py:80-119`` leaves behind: one ``au/<video_id>.txt`` per video (header line, then one
  comma-separated row of twelve 0/1 decisions per processed frame, in sampler order) and ``inference.pkl``
  (``torch.save({'predictions': [n_samples, 21]})`` with the raw model outputs scattered by dataset index).  The decision is the
  reference's ``np.round(sigmoid(logit))`` (half to even, i.e. ``logit > 0``); the hot path already produces it on the device
  (``avf_au_logits_fwd``'s int32 decisions), so nothing but twelve small integers per frame crosses PCIe for the text files.
* ``nearest_interp`` / ``expand_to_video`` / ``postprocess_directory`` — ``postprocess/postprocess.py:29-89``: the challenge wants one
  row per frame of the ORIGINAL video, but only frames with a detected face were processed; every missing frame repeats a
  neighbouring processed row.

Host-side text / pickle writers only: no arithmetic of the model runs here.
"""
from __future__ import annotations

import os
from typing import Callable, Dict, Iterable, List, Optional, Sequence

import numpy as np
import torch

HEADERS = {                                            # test_aff2.py:87-90
    "AU": "AU1,AU2,AU4,AU6,AU7,AU10,AU12,AU15,AU23,AU24,AU25,AU26",
    "VA": "valence,arousal",
    "EX": "Neutral,Anger,Disgust,Fear,Happiness,Sadness,Surprise",
}


def au_row(decisions: Sequence[int]) -> str:
    """Twelve 0/1 decisions as the reference prints them (test_aff2.py:34-36)."""
    if len(decisions) != 12:
        raise ValueError(f"an AU row has 12 entries, got {len(decisions)}")
    return ",".join(str(int(d)) for d in decisions)


def au_decisions(logits: torch.Tensor) -> np.ndarray:
    """np.round(sigmoid(logits[:, :12])) as integers (train.py:155, test_aff2.py:112-113): sigmoid(0) = 0.5 rounds to 0."""
    return (logits[:, :12].detach().float().cpu().numpy() > 0).astype(np.int64)


class AUResultWriter:
    """Collects what the reference's test loop writes.  ``add`` is called per batch in sampler order; a new text file is opened
    (truncating, as the reference's ``open(..., 'w')`` does) whenever the video id differs from the previous row's."""

    def __init__(self, result_path: str, n_samples: int, task: str = "AU"):
        if task != "AU":
            raise NotImplementedError("only the AU task is on the hot path (SURVEY.md section 8); EX / VA writers are not provided")
        self.result_path = result_path
        self.folder = os.path.join(result_path, "au")
        os.makedirs(self.folder, exist_ok=True)
        self.output = torch.zeros((int(n_samples), 21), dtype=torch.float32)
        self._video: Optional[str] = None
        self._fh = None
        self.files: List[str] = []

    def _switch(self, video_id: str) -> None:
        if self._fh is not None:
            self._fh.close()
        self._video = video_id
        path = os.path.join(self.folder, video_id + ".txt")
        self._fh = open(path, "w")
        self._fh.write(HEADERS["AU"] + "\n")
        self.files.append(path)

    def add(self, video_ids: Sequence[str], indices: Iterable[int], out21: torch.Tensor, decisions: Optional[torch.Tensor] = None) -> None:
        """video_ids / indices: one per row of ``out21`` [b, 21] (any device); ``decisions`` [b, 12] = the int32 device decisions
        of the hot path (computed from the logits when absent)."""
        out = out21.detach().float().cpu()
        if out.dim() != 2 or out.shape[1] != 21 or len(video_ids) != out.shape[0]:
            raise ValueError(f"AUResultWriter.add: expected [b, 21] outputs and b video ids, got {tuple(out.shape)} and {len(video_ids)}")
        dec = au_decisions(out) if decisions is None else decisions.detach().cpu().numpy().astype(np.int64)
        idx = torch.as_tensor(list(indices), dtype=torch.long)
        if idx.numel() != out.shape[0] or dec.shape != (out.shape[0], 12):
            raise ValueError("AUResultWriter.add: indices / decisions do not match the batch")
        for r, vid in enumerate(video_ids):
            if vid != self._video:
                self._switch(vid)
            self._fh.write(au_row(dec[r]) + "\n")
        self.output[idx] = out

    def close(self) -> str:
        """Close the open text file and write inference.pkl; returns its path."""
        if self._fh is not None:
            self._fh.close()
            self._fh = None
        path = os.path.join(self.result_path, "inference.pkl")
        torch.save({"predictions": self.output}, path)
        return path

    def __enter__(self):
        return self

    def __exit__(self, *exc):
        self.close()
        return False


def nearest_interp(frame_ids: Sequence[int], target_len: int) -> List[int]:
    """For each frame 0..target_len-1 of the original video, the index (into the SORTED processed frames) of the row to repeat
    (postprocess/postprocess.py:30-48).  Row k covers as many output frames as the gap to the next processed frame,
    counted from output position 0, and the last row fills the rest; with at least as many processed frames as video frames the
    rows are taken one to one.  The list can be longer than target_len when the processed frame numbers run past it — callers
    read the first target_len entries, like the reference."""
    src = sorted(int(f) for f in frame_ids)
    n = len(src)
    if target_len <= n:
        return list(range(n))
    out: List[int] = []
    for k in range(n - 1):
        out.extend([k] * (src[k + 1] - src[k]))
    out.extend([max(n - 1, 0)] * (target_len - len(out)))
    return out


def expand_to_video(pred_lines: Sequence[str], frame_ids: Sequence[int], n_frame: int) -> List[str]:
    """``pred_lines`` = header + one row per processed frame (a file written by AUResultWriter); returns header + n_frame rows
    (postprocess/postprocess.py:77-86)."""
    if len(frame_ids) != len(pred_lines) - 1:
        raise ValueError(f"{len(frame_ids)} processed frames but {len(pred_lines) - 1} prediction rows")
    which = nearest_interp(frame_ids, n_frame)
    return [pred_lines[0]] + [pred_lines[which[i] + 1] for i in range(n_frame)]


def video_file_stem(aligned_name: str) -> str:
    """cropped-aligned folder name -> video file stem (postprocess/postprocess.py:58): subject suffixes dropped."""
    return aligned_name.replace("_main", "").replace("_left", "").replace("_right", "")


def postprocess_directory(prediction_dir: str, out_dir: str, frames_of: Callable[[str], Sequence[int]], n_frames_of: Callable[[str], int]) -> Dict[str, int]:
    """Every ``<prediction_dir>/*.txt`` -> ``<out_dir>/<same name>`` expanded to the original video's frame count.
    ``frames_of(aligned_name)`` = the processed frame numbers (the reference lists ``cropped_aligned/<name>/*.jpg``),
    ``n_frames_of(video_stem)`` = the video's frame count (the reference's ``n_video_frames.pkl``).  Returns rows written per file."""
    os.makedirs(out_dir, exist_ok=True)
    written = {}
    for name in sorted(os.listdir(prediction_dir)):
        if not name.endswith(".txt"):
            continue
        aligned = name[:-4].split(".")[0]
        with open(os.path.join(prediction_dir, name)) as f:
            lines = f.readlines()
        rows = expand_to_video(lines, frames_of(aligned), int(n_frames_of(video_file_stem(aligned))))
        with open(os.path.join(out_dir, name), "w") as f:
            f.writelines(rows)
        written[name] = len(rows) - 1
    return written
