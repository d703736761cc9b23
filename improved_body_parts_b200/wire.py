"""Wire records: what the grouping path hands to its consumer -- the payload of ``format_results``.

The reference turns ``process()``'s per-image person lists into COCO result records (``evaluate.py:563-582``):

    {"image_id": id, "category_id": 1, "keypoints": [x0, y0, v0, ..., x16, y16, v16], "score": s}

with ``v = 1 if x > 0 or y > 0 else 0`` and ``s = 1 - 1/total`` (``:541``).  The assemble kernel writes exactly that
payload per image as one fixed-stride binary record (include/spgroup.h "wire records"):

    int32 n_persons | uint32 status | rows[R] of (17 x (x, y) float64, score float64, presence mask uint64)

(bit g of the mask clear: joint g was not found and the reference stores the INTEGER placeholder ``X, Y = 0, 0``,
evaluate.py:531 -- kept so that the JSON text is the reference's, ``0`` not ``0.0``.)

Only the first ``n_persons`` rows are written.  This module is the host-side view of those records: the numpy dtype,
unpacking into the reference's Python structures, and ``format_results`` itself (same JSON as the reference's).
No arithmetic happens here -- ``v`` is the only derived field, computed exactly as the reference does.
"""
from __future__ import annotations

import json
from typing import Dict, Iterable, List, Sequence, Tuple

import numpy as np

HEADER_BYTES = 8


def record_dtype(n_joints: int = 17, rows: int = 64) -> np.dtype:
    """Structured dtype of one image's record; ``itemsize == spg_wire_record_bytes()``."""
    row = np.dtype([("xy", "<f8", (int(n_joints), 2)), ("score", "<f8"), ("present", "<u8")])
    return np.dtype([("n_persons", "<i4"), ("status", "<u4"), ("rows", row, (int(rows),))])


def record_bytes(n_joints: int = 17, rows: int = 64) -> int:
    return HEADER_BYTES + int(rows) * (2 * int(n_joints) + 2) * 8


def as_records(buf, n_joints: int = 17, rows: int = 64) -> np.ndarray:
    """View a byte buffer (numpy uint8 array / bytes / CPU torch tensor) as an array of records."""
    if hasattr(buf, "numpy") and not isinstance(buf, np.ndarray):
        buf = buf.numpy()
    a = np.ascontiguousarray(np.asarray(buf)).reshape(-1).view(np.uint8)
    dt = record_dtype(n_joints, rows)
    if a.size % dt.itemsize:
        raise ValueError(f"{a.size} bytes is not a whole number of {dt.itemsize}-byte records")
    return a.view(dt)


def pack(people_per_image: Sequence[Iterable], n_joints: int = 17, rows: int = 64, status: Sequence[int] = ()) -> np.ndarray:
    """Host-side producer of records (tests, CPU-side tools): ``process()``-style people lists -> record array.

    A joint given as the reference's integer placeholder ``(0, 0)`` (evaluate.py:531) is stored as absent."""
    out = np.zeros((len(people_per_image),), record_dtype(n_joints, rows))
    for i, people in enumerate(people_per_image):
        people = list(people)
        if len(people) > rows:
            raise ValueError(f"image {i}: {len(people)} persons do not fit {rows} rows")
        out[i]["n_persons"] = len(people)
        if len(status):
            out[i]["status"] = status[i]
        for p, (pts, score) in enumerate(people):
            present = 0
            for g, (x, y) in enumerate(pts):
                if isinstance(x, (int, np.integer)) and isinstance(y, (int, np.integer)) and x == 0 and y == 0:
                    continue
                out[i]["rows"][p]["xy"][g] = (x, y)
                present |= 1 << g
            out[i]["rows"][p]["score"] = score
            out[i]["rows"][p]["present"] = present
    return out


def people_of(rec) -> List[Tuple[List[Tuple[float, float]], float]]:
    """One record -> ``process()``'s return value for that image (evaluate.py:523-543): [([17 x (x, y)], score)]."""
    out = []
    for r in rec["rows"][:int(rec["n_persons"])]:
        present = int(r["present"])
        pts = [(np.float64(x), np.float64(y)) if (present >> g) & 1 else (0, 0) for g, (x, y) in enumerate(r["xy"])]
        out.append((pts, np.float64(r["score"])))
    return out


def unpack(records: np.ndarray, image_ids: Sequence) -> Dict[object, list]:
    """``predict_many``'s dict (evaluate.py:550-560): image id -> people."""
    if len(records) != len(image_ids):
        raise ValueError("one image id per record expected")
    return {iid: people_of(rec) for iid, rec in zip(image_ids, records)}


def coco_results(keypoints: Dict[object, Iterable]) -> List[dict]:
    """The list ``format_results`` dumps (evaluate.py:563-580), value for value."""
    out = []
    for image_id, people in keypoints.items():
        for keypoint_list, score in people:
            flat = []
            for x, y in keypoint_list:
                for v in [x, y, 1 if x > 0 or y > 0 else 0]:
                    flat.append(v)
            out.append({"image_id": image_id, "category_id": 1, "keypoints": flat, "score": score})
    return out


def format_results(keypoints: Dict[object, Iterable], resFile: str) -> None:
    """Drop-in for ``evaluate.format_results`` (evaluate.py:563-582): same arguments, same file contents."""
    json.dump(coco_results(keypoints), open(resFile, "w"))


def coco_results_from_records(records: np.ndarray, image_ids: Sequence) -> List[dict]:
    """Records straight to the list ``format_results`` dumps."""
    return coco_results(unpack(records, image_ids))
