"""Skeleton definition and grouping hyper-parameters consumed by the hot path.

Data only.  The 18-part / 30-limb "Canonical" skeleton is the one
``/root/reference/config/config.py:52-126`` builds (limb table asserted at
``:87-92``; channel layout ``:96-103``; COCO re-ordering ``:117-118``); the
default parameters are the ``[param]`` block of ``/root/reference/utils/config:17-28``.
The limb table is runtime data everywhere in this repo (the reference also
ships 24- and 49-limb skeletons, ``config/config2.py`` / ``config_dense.py``).
"""
from __future__ import annotations

import dataclasses
import re
from typing import Dict, List, Optional, Sequence, Tuple

PART_NAMES: Tuple[str, ...] = (
    "nose", "neck", "Rsho", "Relb", "Rwri", "Lsho", "Lelb", "Lwri", "Rhip", "Rkne",
    "Rank", "Lhip", "Lkne", "Lank", "Reye", "Leye", "Rear", "Lear",
)
NUM_PARTS = len(PART_NAMES)  # 18 keypoint channels used by find_peaks (evaluate.py:175,187)

_LIMB_NAMES: Tuple[Tuple[str, str], ...] = (
    ("neck", "nose"), ("neck", "Reye"), ("neck", "Leye"), ("neck", "Rear"), ("neck", "Lear"),
    ("nose", "Reye"), ("nose", "Leye"), ("Reye", "Rear"), ("Leye", "Lear"),
    ("neck", "Rsho"), ("Rsho", "Relb"), ("Relb", "Rwri"),
    ("neck", "Lsho"), ("Lsho", "Lelb"), ("Lelb", "Lwri"),
    ("neck", "Rhip"), ("Rhip", "Rkne"), ("Rkne", "Rank"),
    ("neck", "Lhip"), ("Lhip", "Lkne"), ("Lkne", "Lank"),
    ("nose", "Rsho"), ("nose", "Lsho"), ("Rsho", "Rhip"), ("Rhip", "Lkne"),
    ("Lsho", "Lhip"), ("Lhip", "Rkne"), ("Rear", "Rsho"), ("Lear", "Lsho"), ("Rhip", "Lhip"),
)
_IDX = {n: i for i, n in enumerate(PART_NAMES)}
#: limbs_conn of the Canonical config: (from_part, to_part) per body-part channel.
LIMBS: Tuple[Tuple[int, int], ...] = tuple((_IDX[a], _IDX[b]) for a, b in _LIMB_NAMES)
NUM_LIMBS = len(LIMBS)  # 30

# pinned exactly as config/config.py:87-92 asserts it
assert [a for a, _ in LIMBS] == [1, 1, 1, 1, 1, 0, 0, 14, 15, 1, 2, 3, 1, 5, 6, 1, 8, 9, 1, 11, 12,
                                 0, 0, 2, 8, 5, 11, 16, 17, 8]
assert [b for _, b in LIMBS] == [0, 14, 15, 16, 17, 14, 15, 16, 17, 2, 3, 4, 5, 6, 7, 8, 9, 10, 11, 12, 13,
                                 2, 5, 8, 12, 11, 9, 2, 5, 11]

#: the 24-limb skeleton of config/config2.py:69-83 (asserted there at :80-83): the limb table is runtime data
LIMBS_24: Tuple[Tuple[int, int], ...] = tuple(zip(
    (1, 1, 1, 1, 1, 0, 0, 14, 15, 1, 2, 3, 1, 5, 6, 1, 8, 9, 1, 11, 12, 8, 2, 5),
    (0, 14, 15, 16, 17, 14, 15, 16, 17, 2, 3, 4, 5, 6, 7, 8, 9, 10, 11, 12, 13, 11, 16, 17)))

#: network channel layout (config/config.py:96-103): body parts first, then keypoints, then 2 background maps.
PAF_START, HEAT_START, BKG_START, NUM_LAYERS = 0, NUM_LIMBS, NUM_LIMBS + NUM_PARTS, NUM_LIMBS + NUM_PARTS + 2

#: detection part index -> COCO keypoint index (None: neck is dropped), config/config.py:117-118
DT_GT_MAPPING: Dict[int, Optional[int]] = {0: 0, 1: None, 2: 6, 3: 8, 4: 10, 5: 5, 6: 7, 7: 9, 8: 12, 9: 14,
                                           10: 16, 11: 11, 12: 13, 13: 15, 14: 2, 15: 1, 16: 4, 17: 3}
#: COCO keypoint index -> detection part index (17 entries), the inverse used on the device
COCO_FROM_PART: Tuple[int, ...] = tuple(
    next(dt for dt, gt in DT_GT_MAPPING.items() if gt == g) for g in range(17))

#: flip-ensemble channel permutations (config/config.py:121-124); used by predict(), a "next" row
FLIP_HEAT_ORD = (0, 1, 5, 6, 7, 2, 3, 4, 11, 12, 13, 8, 9, 10, 15, 14, 17, 16, 18, 19)
FLIP_PAF_ORD = (0, 2, 1, 4, 3, 6, 5, 8, 7, 12, 13, 14, 9, 10, 11, 18, 19, 20, 15, 16, 17, 22, 21, 25, 26,
                23, 24, 28, 27, 29)


@dataclasses.dataclass
class GroupParams:
    """Grouping hyper-parameters; names are the reference's dict keys (utils/config:17-28).

    ``min_parts`` / ``min_mean_score`` are the two literals of the final prune at
    evaluate.py:493 (demo_image.py:533 uses 4 instead of 2).
    """

    thre1: float = 0.1
    thre2: float = 0.1
    connect_ration: float = 0.8
    mid_num: int = 20
    len_rate: float = 16.0
    connection_tole: float = 0.7
    offset_radius: int = 2
    remove_recon: int = 0
    min_parts: int = 2
    min_mean_score: float = 0.45
    #: demo_image.py's inlined copy of the grouping code: ``>`` at :288 where evaluate.py:246 has ``>=``,
    #: and a limb-length check in the same-B refresh (:414-415).  0 = evaluate.py.
    crit1_strict: int = 0
    refresh_len_check: int = 0

    @classmethod
    def from_dict(cls, params: dict) -> "GroupParams":
        kw = {}
        for f in dataclasses.fields(cls):
            if f.name in params:
                kw[f.name] = type(f.default)(params[f.name])
        return cls(**kw)

    def to_dict(self) -> dict:
        d = dataclasses.asdict(self)
        for k in ("min_parts", "min_mean_score", "crit1_strict", "refresh_len_check"):
            d.pop(k)
        return d

    @classmethod
    def demo(cls, params: Optional[dict] = None) -> "GroupParams":
        """The behaviour of demo_image.py's inlined grouping (``:288``, ``:414-415``, ``:533``) for a params dict."""
        gp = cls.from_dict(dict(params or {}))
        gp.crit1_strict, gp.refresh_len_check, gp.min_parts = 1, 1, 4
        return gp


def default_params() -> dict:
    """The dict ``config_reader()`` would hand to the grouping functions (hot-path keys only)."""
    return GroupParams().to_dict()


def read_reference_ini(path: str) -> Tuple[dict, dict]:
    """Minimal reader for the reference's ``utils/config`` INI (configobj is not installed).

    Mirrors the conversions of ``utils/config_reader.py:6-37`` for the keys it converts, including
    the quirk that ``scale_search = 1`` becomes ``[1.0]`` by iterating the *characters* of ``'1'`` (:22).
    Returns ``(param, model)`` like ``config_reader()``.
    """
    param: dict = {}
    models: Dict[str, dict] = {}
    section: Optional[dict] = None
    with open(path, encoding="utf-8") as fh:
        for raw in fh:
            line = raw.split("#", 1)[0].strip()
            if not line:
                continue
            m2 = re.fullmatch(r"\[\[(.+)\]\]", line)
            m1 = re.fullmatch(r"\[(.+)\]", line)
            if m2:
                section = models.setdefault(m2.group(1).strip(), {})
            elif m1:
                section = param if m1.group(1).strip() == "param" else {}
            elif "=" in line and section is not None:
                k, v = (s.strip() for s in line.split("=", 1))
                section[k] = v.strip("'\"")
    model = dict(models[param["modelID"]])
    for k in ("boxsize", "stride", "max_downsample", "padValue"):
        model[k] = int(model[k])
    for k in ("remove_recon", "use_gpu", "mid_num", "min_num", "offset_radius", "GPUdeviceNumber"):
        param[k] = int(param[k])
    for k in ("starting_range", "ending_range", "thre1", "thre2", "connect_ration", "connection_tole",
              "len_rate", "crop_ratio", "bbox_ratio"):
        param[k] = float(param[k])
    param["scale_search"] = list(map(float, param["scale_search"]))
    param["rotation_search"] = list(map(float, param["rotation_search"]))
    return param, model
