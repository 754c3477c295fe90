"""CPU oracle (TEST INFRASTRUCTURE ONLY) of the planner's grasp post-processing: `process` and `select` of
src/nr/main.py:23-74, restated with the same scipy.ndimage calls (scipy is the reference's own dependency for this step,
so the arithmetic IS the reference's).  select() returns index / score / rot / width rows instead of gd.grasp.Grasp objects
(main.py:77-84 only wraps them).  Pinned to the reference's source by tests/test_grasp_post.py where /root/reference exists."""
import numpy as np
from scipy import ndimage


def process(tsdf_vol, qual_vol, rot_vol, width_vol, gaussian_filter_sigma=1.0, min_width=1.33, max_width=9.33,
            tsdf_thres_high=0.5, tsdf_thres_low=1e-3):
    """main.py:23-55."""
    tsdf_vol, qual_vol, rot_vol, width_vol = tsdf_vol.squeeze(), qual_vol.squeeze(), rot_vol.squeeze(), width_vol.squeeze()
    qual_vol = ndimage.gaussian_filter(qual_vol, sigma=gaussian_filter_sigma, mode='nearest')            # main.py:39-41
    outside = tsdf_vol > tsdf_thres_high                                                                   # main.py:44
    inside = np.logical_and(tsdf_thres_low < tsdf_vol, tsdf_vol < tsdf_thres_high)                         # main.py:45
    valid = ndimage.binary_dilation(outside, iterations=2, mask=np.logical_not(inside))                    # main.py:46-48
    qual_vol[valid == False] = 0.0                                                                         # noqa: E712  main.py:49
    qual_vol[np.logical_or(width_vol < min_width, width_vol > max_width)] = 0.0                            # main.py:52
    return qual_vol, rot_vol, width_vol


def select(qual_vol, rot_vol, width_vol, threshold=0.90, max_filter_size=4):
    """main.py:58-74; returns (indices [G,3] int, scores [G], rots [G,4], widths [G]) in np.argwhere order."""
    qual_vol = qual_vol.copy()
    qual_vol[qual_vol < threshold] = 0.0
    max_vol = ndimage.maximum_filter(qual_vol, size=max_filter_size)
    qual_vol = np.where(qual_vol == max_vol, qual_vol, 0.0)
    idx = np.argwhere(np.where(qual_vol, 1.0, 0.0))
    i, j, k = idx[:, 0], idx[:, 1], idx[:, 2]
    return idx, qual_vol[i, j, k], rot_vol[:, i, j, k].T, width_vol[i, j, k]
