"""cv::FREAK pin (reference call site MoFREAKUtilities.cpp:427-428).  The oracle's FREAK is a restatement of OpenCV 2.4.2's
freak.cpp made without access to any FREAK implementation (tests/golden/FREAK_SEARCH.md): PARITY UNPINNED until the vectors
of tests/golden/make_freak_vectors.py exist.  With them, this test is the pin: same survivors, same order, same 64 bytes."""
import glob
import os

import numpy as np
import pytest

from tests import util
from tests.golden import make_freak_vectors as mk

VECTORS = sorted(glob.glob(os.path.join(util.GOLDEN, "freak_*.npz")))


def test_freak_vectors_script_is_importable_and_reports_availability():
    ex, how = mk.freak_extractor()
    assert isinstance(how, str) and how
    if ex is None:
        assert "FREAK" in how or "cv2" in how


@pytest.mark.skipif(not VECTORS, reason="FREAK PARITY UNPINNED: no OpenCV FREAK vectors (tests/golden/freak_*.npz); run "
                                        "tests/golden/make_freak_vectors.py where opencv-contrib is installed")
@pytest.mark.parametrize("path", VECTORS or ["none"])
def test_oracle_freak_equals_opencv_vectors(orc, path):
    name = os.path.basename(path)[len("freak_"):-len(".npz")]
    diff, kps = mk.inputs(name)
    want = np.load(path)
    got_kp, got_desc, _ = orc.freak_compute(diff, kps)
    assert got_kp.shape[0] == want["keypoints"].shape[0], "FREAK's border filter keeps a different keypoint set"
    assert np.array_equal(got_kp[:, :3].astype(np.float32).view(np.uint32), want["keypoints"].view(np.uint32))
    assert np.array_equal(got_desc, want["descriptors"]), "descriptor bytes differ from OpenCV (%s)" % str(want["how"])


def test_inputs_of_the_pin_are_reproducible(orc):
    """The frames and keypoints the vectors are made from come from the synthetic generator and the reference's own
    detector, so a later run of the script sees exactly the inputs this container sees."""
    diff, kps = mk.inputs("odd_212x163x9_seed7")
    assert diff.shape == (163, 212) and kps.shape[1] == 5 and len(kps) > 10
    kp2, desc, theta = orc.freak_compute(diff, kps)
    assert len(kp2) <= len(kps) and desc.shape == (len(kp2), 64) and ((theta >= 0) & (theta < 256)).all()
