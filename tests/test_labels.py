"""Row N4, label half, on the device (-m gpu): dvp_label_segment == EdgeSegment(scale, image, 1) (reference APD.cpp:348-402,
437-499).  Checked against the golden label maps made by the real OpenCV 4.13 (tests/golden/label_segment.npz) and against
the pinned CPU restatement (oracle/cpu/label_cpu.cpp) on renders up to 1555 x 1037.  The quarter-size edge image after the
Hough lines must match byte for byte (which pins resize, Roberts, region borders, HoughLinesP in cv::RNG order and cv::line);
the label map must be the same PARTITION with the same classes (0 boundary, -1 small region, > 0 region) — the hot path
compares labels only for equality (APD.cu:3461, 3629, 3857-3886), and the ids are the library's own."""
import os

import numpy as np
import pytest

from util import GOLDEN
from test_labels_oracle import segment as segment_oracle

pytestmark = pytest.mark.gpu


def same_partition(a: np.ndarray, b: np.ndarray):
    """Same classes (0 / -1 / positive) everywhere and a one-to-one correspondence between the positive ids."""
    if a.shape != b.shape or not ((a == 0) == (b == 0)).all() or not ((a == -1) == (b == -1)).all():
        return False
    pos = a > 0
    if not pos.any():
        return True
    pairs = np.unique(np.stack([a[pos], b[pos]], 1), axis=0)
    return len(np.unique(pairs[:, 0])) == len(pairs) == len(np.unique(pairs[:, 1]))


def test_gpu_label_segment_matches_opencv_golden_vectors():
    from dvp_mvs_b200 import label_segment
    g = np.load(os.path.join(GOLDEN, "label_segment.npz"))
    for i in range(int(g["count"])):
        img = g[f"image_{int(g[f'image_of_{i}'])}"]
        labels, small, ms = label_segment(img, int(g[f"scale_{i}"]))
        np.testing.assert_array_equal(small, g[f"edge_small_{i}"], err_msg=f"case {i}: edge image after the Hough lines")
        assert same_partition(labels, g[f"labels_{i}"]), i


@pytest.mark.parametrize("size,scale", [((640, 480), 1), ((640, 480), 2), ((1555, 1037), 1), ((1555, 1037), 3), ((333, 211), 1), ((64, 48), 1)])
def test_gpu_label_segment_matches_the_restatement_on_renders(size, scale):
    from dvp_mvs_b200 import label_segment, synth
    W, H = size
    sc = synth.make_scene(W, H, 1)
    img = np.clip(np.rint(sc.images[0]), 0, 255).astype(np.uint8)
    want, want_small = segment_oracle(img, scale)
    labels, small, ms = label_segment(img, scale)
    np.testing.assert_array_equal(small, want_small)
    assert same_partition(labels, want)
    assert ((labels > 0).sum() > 0) or W < 100          # the renders have large textureless regions


def test_gpu_label_segment_rejects_bad_arguments():
    from dvp_mvs_b200 import label_segment, DvpError
    with pytest.raises(DvpError):
        label_segment(np.zeros((8, 8), np.uint8), 1)       # below 16 x 16
    with pytest.raises(DvpError):
        label_segment(np.zeros((64, 64), np.uint8), 9)
