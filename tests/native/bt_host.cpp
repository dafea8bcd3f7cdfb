// TEST BUILD ONLY — the K7 ByteTrack core (hockey-vision-analytics_b200/csrc/k7_bytetrack_core.h) compiled for the host
// with g++ (BT_NL = 1: the warp-cooperative statements run serially), so that tests/test_bytetrack_core.py can compare
// the tracker's logic with oracle/bytetrack_restated.py and its assignment solver with scipy on a machine without a
// GPU.  Built by tests/native/build.py into tests/native/_build/ (git-ignored); the product never loads it.
#include <stdlib.h>
#include <string.h>

#include "k7_bytetrack_core.h"

struct BtHost {
    BtParams p;
    BtClip clip;
    BtWork work;
    double* cost;
};

extern "C" {

void* bt_host_create(double act_thr, double det_thr, double match_thr, int max_time_lost, int min_consec) {
    BtHost* h = (BtHost*)calloc(1, sizeof(BtHost));
    h->p.match_thr = match_thr; h->p.det_thr = det_thr; h->p.act_thr = (float)act_thr;
    h->p.max_time_lost = max_time_lost; h->p.min_consec = min_consec;
    h->p.min_conf = -INFINITY; h->p.class_mask = 0xFFFFFFFFu;
    h->cost = (double*)malloc(sizeof(double) * 2 * BT_N * BT_N);
    bt_reset(&h->clip);
    return h;
}

void bt_host_destroy(void* hv) { BtHost* h = (BtHost*)hv; free(h->cost); free(h); }

void bt_host_reset(void* hv) { bt_reset(&((BtHost*)hv)->clip); }

// one frame; returns the number of kept detections (or -1), rows / ids in out_row / out_tid
int bt_host_update(void* hv, const float* xyxy, const float* conf, const int32_t* cls, int n, float min_conf,
                   uint32_t class_mask, int32_t* out_row, int32_t* out_tid) {
    BtHost* h = (BtHost*)hv;
    BtParams p = h->p;
    p.min_conf = min_conf; p.class_mask = class_mask;
    return bt_update(&h->clip, &h->work, p, xyxy, conf, cls, n, h->cost, h->cost + BT_N * BT_N, out_row, out_tid);
}

// the assignment solver alone: cost [na, nb] row-major -> matches (count returned), unmatched rows / columns
int bt_host_assign(const double* cost_in, int na, int nb, double thresh, int32_t* ma, int32_t* mb, int32_t* n_ua, int32_t* n_ub) {
    BtWork* w = (BtWork*)calloc(1, sizeof(BtWork));
    double* c = (double*)malloc(sizeof(double) * 2 * BT_N * BT_N);
    memcpy(c, cost_in, sizeof(double) * na * nb);
    bt_assign(w, na, nb, c, c + BT_N * BT_N, thresh);
    const int nm = w->n_match;
    for (int k = 0; k < nm; k++) { ma[k] = w->ma[k]; mb[k] = w->mb[k]; }
    *n_ua = w->n_ua; *n_ub = w->n_ub;
    free(c); free(w);
    return nm;
}

// raw linear_sum_assignment (no clamping): col4row for nr <= nc
int bt_host_lsap(const double* cost, int nr, int nc, int32_t* col4row) {
    BtWork* w = (BtWork*)calloc(1, sizeof(BtWork));
    const int rc = bt_lsap(nr, nc, cost, w);
    for (int i = 0; i < nr; i++) col4row[i] = w->col4row[i];
    free(w);
    return rc;
}

}  // extern "C"
