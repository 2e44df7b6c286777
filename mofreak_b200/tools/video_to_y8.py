#!/usr/bin/env python
"""Frame ingest outside the hot path: decode a video and convert BGR -> gray exactly as the reference does
(cv::VideoCapture + cv::cvtColor(CV_BGR2GRAY), MoFREAKUtilities.cpp:380-410), writing the raw MFKGRAY1 container that
MoFREAKUtilities::computeMoFREAKFromFile of mofreak_b200/host reads.

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


def main():
    import cv2
    src, dst = sys.argv[1], sys.argv[2]
    cap = cv2.VideoCapture(src)
    if not cap.isOpened():
        raise SystemExit("Could not open file: " + src)
    frames = []
    while True:
        ok, frame = cap.read()
        if not ok:
            break
        frames.append(cv2.cvtColor(frame, cv2.COLOR_BGR2GRAY))
    write_y8(dst, np.stack(frames))
    print("%s: %d frames %dx%d" % (dst, len(frames), frames[0].shape[1], frames[0].shape[0]))


if __name__ == "__main__":
    main()
