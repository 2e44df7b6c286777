"""Self-activating pin for the one unpinned stage, cv::FREAK (reference call site MoFREAKUtilities.cpp:427-428).

The reference links OpenCV 2.4.2's FREAK, which is not vendored; no FREAK implementation exists in the build container
(search recorded in FREAK_SEARCH.md), so oracle/freak_restate.cpp is PARITY UNPINNED.  Run this script on ANY machine that
has one - `cv2.xfeatures2d.FREAK_create` (opencv-contrib-python) or an OpenCV 2.4 `cv2.DescriptorExtractor_create("FREAK")`:

    python tests/golden/make_freak_vectors.py          # writes tests/golden/freak_<case>.npz

For every golden case it takes the LAST processed frame's difference image and the keypoints the reference's own BRISK
detector finds on it (both reproducible here: synthetic clip + oracle/_ref), runs OpenCV's FREAK with its defaults
(orientationNormalized, scaleNormalized, patternScale 22, nOctaves 4 - exactly `cv::FREAK extractor;` of the reference) and
stores the surviving keypoints and their 64-byte descriptors.  tests/test_freak_pin.py then compares
oracle.freak_compute against these files; until they exist it SKIPS, saying that FREAK parity is unpinned.
"""
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)

import numpy as np  # noqa: E402

HERE = os.path.dirname(os.path.abspath(__file__))


def freak_extractor():
    """(extractor, description) from whatever OpenCV is importable, or (None, why)."""
    try:
        import cv2
    except Exception as e:
        return None, "cv2 not importable: %r" % (e,)
    xf = getattr(cv2, "xfeatures2d", None)
    if xf is not None and hasattr(xf, "FREAK_create"):
        return xf.FREAK_create(), "cv2 %s xfeatures2d.FREAK_create() defaults" % cv2.__version__
    if hasattr(cv2, "DescriptorExtractor_create"):                 # OpenCV 2.4 python bindings
        try:
            return cv2.DescriptorExtractor_create("FREAK"), "cv2 %s DescriptorExtractor_create('FREAK')" % cv2.__version__
        except Exception:
            pass
    return None, "cv2 %s has neither xfeatures2d.FREAK_create nor DescriptorExtractor_create('FREAK')" % cv2.__version__


def inputs(name):
    """(difference image, BRISK keypoints [n, 5] = x, y, size, response, layer) of the case's last processed frame."""
    import mofreak_b200
    from tests import oracle_binding
    meta = json.load(open(os.path.join(HERE, "golden.json")))[name]
    clip = mofreak_b200.synth_clip(meta["width"], meta["height"], meta["frames"], meta["seed"])
    c = meta["frames"] - 1
    diff = np.abs(clip[c].astype(np.int16) - clip[c - 5].astype(np.int16)).astype(np.uint8)
    orc = oracle_binding.load()
    return diff, orc.brisk(diff)


def main():
    ex, how = freak_extractor()
    if ex is None:
        raise SystemExit("no FREAK available (%s): nothing written, FREAK parity stays unpinned" % how)
    import cv2
    for name in json.load(open(os.path.join(HERE, "golden.json"))):
        diff, kps = inputs(name)
        cv_kps = [cv2.KeyPoint(float(x), float(y), float(s), -1.0, float(r), int(l)) for x, y, s, r, l in kps]
        cv_kps, desc = ex.compute(diff, cv_kps)
        out = np.array([[k.pt[0], k.pt[1], k.size] for k in cv_kps], np.float32).reshape(-1, 3)
        desc = np.zeros((0, 64), np.uint8) if desc is None else np.asarray(desc, np.uint8)
        np.savez_compressed(os.path.join(HERE, "freak_%s.npz" % name), keypoints=out, descriptors=desc, how=np.array(how))
        print(name, len(kps), "keypoints in,", len(out), "described;", how)


if __name__ == "__main__":
    main()
