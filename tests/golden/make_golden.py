"""Regenerates the golden fixtures.  Run where /root/reference is present, so that the detector inside the oracle
is the reference's OWN object code (oracle/_ref); FREAK/MIP are the oracle's restatements (OpenCV is absent).

    python tests/golden/make_golden.py

Each case: synthetic clip (mofreak_b200/host/synth.cpp) -> oracle rows -> <name>.rows.bin (32-byte mfk_row records,
bit-exact target) and <name>.mofreak (the reference's text format, byte-exact target of the writer)."""
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)

import numpy as np  # noqa: E402

import mofreak_b200  # noqa: E402
from tests import oracle_binding  # noqa: E402

HERE = os.path.dirname(os.path.abspath(__file__))
# name: (width, height, frames, seed)   — BASELINE.json configs[0] first
CASES = {
    "kth_160x120x100_seed1": (160, 120, 100, 1),
    "ucf_320x240x12_seed3": (320, 240, 12, 3),
    "trecvid_720x576x7_seed5": (720, 576, 7, 5),
    "odd_212x163x9_seed7": (212, 163, 9, 7),
}


def main():
    orc = oracle_binding.load()
    if not orc.has_reference:
        raise SystemExit("refusing to write goldens without the reference's detector (oracle/_ref)")
    meta = {}
    for name, (w, h, n, seed) in CASES.items():
        clip = mofreak_b200.synth_clip(w, h, n, seed)
        rows = orc.extract_clip(clip, use_ref=True)
        rows.tofile(os.path.join(HERE, name + ".rows.bin"))
        with open(os.path.join(HERE, name + ".mofreak"), "w") as f:
            f.write(orc.format_mofreak(rows))
        meta[name] = {"width": w, "height": h, "frames": n, "seed": seed, "rows": int(len(rows)),
                      "clip_sha1": __import__("hashlib").sha1(clip.tobytes()).hexdigest()}
        print(name, len(rows), "rows")
    json.dump(meta, open(os.path.join(HERE, "golden.json"), "w"), indent=1, sort_keys=True)


if __name__ == "__main__":
    main()
