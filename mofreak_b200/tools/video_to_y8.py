#!/usr/bin/env python
"""Frame ingest outside the hot path: decode a video and convert BGR -> gray as the reference does
(cv::VideoCapture + cv::cvtColor(CV_BGR2GRAY), MoFREAKUtilities.cpp:380-410), writing the raw MFKGRAY1 container that
MoFREAKUtilities::computeMoFREAKFromFile of mofreak_b200/host reads.

The gray conversion is OpenCV 2.4's (the reference's version): (B*1868 + G*9617 + R*4899 + 8192) >> 14.  OpenCV >= 3
uses 15-bit coefficients and differs in about 0.3 % of the pixels; pass --cv3 for that arithmetic.  The same two forms
are available on the device (mfk_upload_frames_bgr, coeff_bits 14 / 15).

    python mofreak_b200/tools/video_to_y8.py in.avi out.y8
"""
import sys

import numpy as np


def write_y8(path, frames):
    frames = np.ascontiguousarray(frames, np.uint8)
    n, h, w = frames.shape
    with open(path, "wb") as f:
        f.write(b"MFKGRAY1")
        f.write(np.array([w, h, n], "<i4").tobytes())
        f.write(frames.tobytes())


def read_y8(path):
    with open(path, "rb") as f:
        assert f.read(8) == b"MFKGRAY1", "not a MFKGRAY1 file"
        w, h, n = np.frombuffer(f.read(12), "<i4")
        return np.frombuffer(f.read(), np.uint8).reshape(n, h, w)


def bgr_to_gray(bgr, cv3=False):
    cb, cg, cr, bits = (3735, 19235, 9798, 15) if cv3 else (1868, 9617, 4899, 14)
    b, g, r = (bgr[..., i].astype(np.int64) for i in range(3))
    return ((b * cb + g * cg + r * cr + (1 << (bits - 1))) >> bits).astype(np.uint8)


def main():
    import cv2
    args = [a for a in sys.argv[1:] if a != "--cv3"]
    cv3 = "--cv3" in sys.argv
    src, dst = args[0], args[1]
    cap = cv2.VideoCapture(src)
    if not cap.isOpened():
        raise SystemExit("Could not open file: " + src)
    frames = []
    while True:
        ok, frame = cap.read()
        if not ok:
            break
        frames.append(bgr_to_gray(frame, cv3))
    write_y8(dst, np.stack(frames))
    print("%s: %d frames %dx%d" % (dst, len(frames), frames[0].shape[1], frames[0].shape[0]))


if __name__ == "__main__":
    main()
