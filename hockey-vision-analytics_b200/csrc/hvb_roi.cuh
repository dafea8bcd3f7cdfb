// Crop / jersey-ROI geometry shared by the K3 kernels.
#pragma once
#include "hvb_common.cuh"

struct hvb_rect { int top, bottom, left, right; };

// HybridTeamClassifier.extract_jersey_region (team_hybrid.py:49-64) and
// TeamClassifier.extract_jersey_region (team.py:76-99).  Python evaluates int(h * 0.1) as a double
// product truncated toward zero; IEEE double multiplication is identical on the device.
__device__ __forceinline__ hvb_rect hvb_roi_rect(int h, int w, int mode) {
    hvb_rect r{0, h, 0, w};
    if (mode == HVB_ROI_HYBRID) {
        if (h < 40 || w < 20) return r;
        r.top = (int)__dmul_rn((double)h, 0.1);
        r.bottom = (int)__dmul_rn((double)h, 0.6);
        r.left = (int)__dmul_rn((double)w, 0.2);
        r.right = (int)__dmul_rn((double)w, 0.8);
    } else if (mode == HVB_ROI_SIMPLE) {
        if (h < 30 || w < 20) return r;
        hvb_rect s;
        s.top = (int)__dmul_rn((double)h, 0.25);
        s.bottom = (int)__dmul_rn((double)h, 0.75);
        s.left = (int)__dmul_rn((double)w, 0.3);
        s.right = (int)__dmul_rn((double)w, 0.7);
        if ((s.bottom - s.top) * (s.right - s.left) == 0) return r;   // empty slice -> whole crop
        return s;
    } else if (mode == HVB_ROI_SEGMENT) {
        // fallback mask of SegmentationTeamClassifier.segment_player (team_segmentation.py:87-96); an empty
        // rectangle stays empty (the reference then reports its "not enough pixels" defaults)
        r.top = (int)__dmul_rn((double)h, 0.2);
        r.bottom = (int)__dmul_rn((double)h, 0.6);
        r.left = (int)__dmul_rn((double)w, 0.3);
        r.right = (int)__dmul_rn((double)w, 0.7);
    }
    return r;
}
