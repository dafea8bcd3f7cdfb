// K7 — ByteTrack, one warp per clip: the whole of sv.ByteTrack.update_with_detections as the reference drives it
// (hockey/main.py:162-168, 207-211 construct it; :228 and :265 step it once per frame).  SURVEY.md §8(f) rank 1.
//
// This header is the tracker itself, written once for two compilers:
//   * nvcc (k7_bytetrack.cu): BT_NL = 32 lanes of one warp cooperate — the Kalman predict / update run one track per
//     lane, the IoU cost matrices one pair per lane, the inner loop of the assignment solver one column per lane —
//     and lane 0 does the list bookkeeping; __syncwarp() orders the phases.
//   * g++ (tests/native/bt_host.cpp, TEST BUILD ONLY): BT_NL = 1, the same statements run serially, so the logic can be
//     compared with the restated supervision tracker on a machine without a GPU.  The product never links that build.
//
// Semantics follow hvb/tracker.py (the host drop-in this kernel replaces) statement by statement, which in turn equals
// oracle/bytetrack_restated.py on ids and kept detections: three association rounds (high-score detections fused with
// the score at minimum_matching_threshold, low-score at 0.5, unconfirmed at 0.7), births above det_thresh, lost tracks
// dropped after max_time_lost, duplicate removal between tracked and lost at IoU distance 0.15, external ids issued
// after minimum_consecutive_frames, and the final detection -> track assignment at 0.5 that puts tracker ids on the
// detections.  The assignment solver is the shortest-augmenting-path algorithm of scipy.optimize.linear_sum_assignment
// (Crouse 2016) with the same tie-breaking, so equal-cost choices come out as scipy makes them.
// All floating-point expressions are written in numpy's evaluation order; the library is built with --fmad=false (the
// test build with -ffp-contract=off), so they round like numpy's.
#pragma once

#include <math.h>
#include <stdint.h>

#if defined(__CUDACC__)
#define BT_HD __host__ __device__ inline
#else
#define BT_HD inline
#endif
#if defined(__CUDA_ARCH__)
#define BT_LANE ((int)(threadIdx.x & 31))
#define BT_NL 32
#define BT_SYNC() __syncwarp()
#else
#define BT_LANE 0
#define BT_NL 1
#define BT_SYNC() ((void)0)
#endif

constexpr int BT_T = 256;          // track slots per clip (tracked + lost + just-removed)
constexpr int BT_D = 320;          // detections per frame (K2a's max_det is 300)
constexpr int BT_N = BT_D;         // max(BT_T, BT_D): side length of the cost matrices
enum { BT_NEW = 0, BT_TRACKED = 1, BT_LOST = 2, BT_REMOVED = 3 };

struct BtParams {
    double match_thr;              // minimum_matching_threshold (first association)
    double det_thr;                // track_activation_threshold + 0.1
    float act_thr;                 // track_activation_threshold
    int max_time_lost;             // int(frame_rate / 30 * lost_track_buffer)
    int min_consec;                // minimum_consecutive_frames
    float min_conf;                // detections enter the tracker iff conf > min_conf ...
    uint32_t class_mask;           // ... and bit class_id of this mask is set (main.py:189-193)
};

struct BtClip {                    // persistent state of one clip's tracker (device global memory)
    int frame_id;
    int next_external;
    int n_tracked, n_lost, n_removed;
    int overflow;                  // sticky: a capacity (track slots / detections per frame) was exceeded
    int next_seq;                  // sequence number of the next chunk this clip accepts (chunks are stepped strictly in order)
    short tracked[BT_T], lost[BT_T], removed[BT_T];      // ordered like supervision's lists
    unsigned char state[BT_T], activated[BT_T], in_use[BT_T];
    int frame[BT_T], start[BT_T], len[BT_T], ext[BT_T];
    double mean[BT_T][8];
    double cov[BT_T][64];
};

struct BtWork {                    // per-frame scratch of one clip (shared memory on the device)
    int n0, n_hi, n_lo;
    short row0[BT_D];              // K2a row of filtered detection i
    short hi[BT_D], lo[BT_D];      // filtered-detection indices of the high / low score sets
    float xyxy[BT_D][4], tlwh[BT_D][4], box[BT_D][4], score[BT_D];
    int n_pool, n_unc, n_rest, n_rem, n_out;
    short pool[BT_T], unc[BT_T], rest[BT_T], rem[BT_D], out[BT_T];
    int n_act, n_refind, n_newlost, n_newrem;
    short act[BT_T], refind[BT_T], newlost[BT_T], newrem[BT_T];
    int n_match, n_ua, n_ub;
    short ma[BT_N], mb[BT_N], ua[BT_N], ub[BT_N];
    int ids[BT_D];
    double u[BT_N], v[BT_N], spc[BT_N];
    int path[BT_N], col4row[BT_N], row4col[BT_N], remaining[BT_N];
    unsigned char SR[BT_N], SC[BT_N];
    double abox[BT_N][4], bbox[BT_N][4], bscore[BT_N];
    unsigned char was[BT_T];
    int fail;
};

// ------------------------------------------------------------------------------------------------ small helpers
BT_HD void bt_tlbr(const double* mean, double* o) {        // STrack.tlbr: xyah -> tlwh -> tlbr
    double m0 = mean[0], m1 = mean[1], m2 = mean[2], m3 = mean[3];
    m2 = m2 * m3;
    m0 = m0 - m2 / 2;
    m1 = m1 - m3 / 2;
    m2 = m2 + m0;
    m3 = m3 + m1;
    o[0] = m0; o[1] = m1; o[2] = m2; o[3] = m3;
}

// 1 - IoU (+ fuse_score), bit for bit numpy's box_iou_batch incl. the float32-area quirk (flags bit 0: a is a float32
// array, bit 1: b is) — the same expression as K4b's iou_cost_kernel.
BT_HD double bt_iou_cost(const double* a, const double* b, int flags, const double* score) {
    const double ax1 = a[0], ay1 = a[1], ax2 = a[2], ay2 = a[3];
    const double bx1 = b[0], by1 = b[1], bx2 = b[2], by2 = b[3];
    double area_a, area_b;
    if (flags & 1) { float w = (float)ax2 - (float)ax1, h = (float)ay2 - (float)ay1; float p = w * h; area_a = (double)p; }
    else area_a = (ax2 - ax1) * (ay2 - ay1);
    if (flags & 2) { float w = (float)bx2 - (float)bx1, h = (float)by2 - (float)by1; float p = w * h; area_b = (double)p; }
    else area_b = (bx2 - bx1) * (by2 - by1);
    const double w = fmax(fmin(ax2, bx2) - fmax(ax1, bx1), 0.0);
    const double h = fmax(fmin(ay2, by2) - fmax(ay1, by1), 0.0);
    const double inter = w * h;
    double iou = inter / ((area_a + area_b) - inter);
    if (isnan(iou)) iou = 0.0;                                       // np.nan_to_num
    else if (isinf(iou)) iou = iou > 0 ? 1.7976931348623157e308 : -1.7976931348623157e308;
    double cost = 1.0 - iou;
    if (score) cost = 1.0 - (1.0 - cost) * (*score);
    return cost;
}

// cost[i * nb + j] for the boxes staged in w->abox / w->bbox (+ w->bscore when fuse)
BT_HD void bt_cost_matrix(BtWork* w, int na, int nb, int flags, bool fuse, double* cost) {
    const int total = na * nb;
    for (int p = BT_LANE; p < total; p += BT_NL) {
        const int i = p / nb, j = p - i * nb;
        cost[p] = bt_iou_cost(w->abox[i], w->bbox[j], flags, fuse ? &w->bscore[j] : nullptr);
    }
    BT_SYNC();
}

// ------------------------------------------------------------------------------------------------ Kalman filter (xyah)
constexpr double BT_W_POS = 1.0 / 20, BT_W_VEL = 1.0 / 160;

BT_HD void bt_predict(BtClip* c, int t) {                 // KalmanFilter.multi_predict for one track
    double* m = c->mean[t];
    double* P = c->cov[t];
    if (c->state[t] != BT_TRACKED) m[7] = 0;
    const double h = m[3];
    const double sp = BT_W_POS * h, sv = BT_W_VEL * h;
    const double q[8] = {sp * sp, sp * sp, 1e-2 * 1e-2, sp * sp, sv * sv, sv * sv, 1e-5 * 1e-5, sv * sv};
    for (int i = 0; i < 4; i++) m[i] = m[i] + m[i + 4];
    // (F @ P) @ F^T + Q with F = I + shift: rows first, then columns, like numpy's left-to-right matmul
    for (int j = 0; j < 8; j++)
        for (int i = 0; i < 4; i++) P[i * 8 + j] = P[i * 8 + j] + P[(i + 4) * 8 + j];
    for (int i = 0; i < 8; i++)
        for (int j = 0; j < 4; j++) P[i * 8 + j] = P[i * 8 + j] + P[i * 8 + j + 4];
    for (int i = 0; i < 8; i++) P[i * 8 + i] = P[i * 8 + i] + q[i];
}

BT_HD void bt_correct(BtClip* c, int t, const float* z32) {   // KalmanFilter.update for one track, measurement xyah
    double* m = c->mean[t];
    double* P = c->cov[t];
    const double h = m[3];
    const double sp = BT_W_POS * h;
    const double r[4] = {sp * sp, sp * sp, 1e-1 * 1e-1, sp * sp};
    double S[16], L[16];
    for (int i = 0; i < 4; i++)
        for (int j = 0; j < 4; j++) S[i * 4 + j] = P[i * 8 + j] + (i == j ? r[i] : 0.0);
    // Cholesky S = L L^T (lower), as scipy.linalg.cho_factor
    for (int i = 0; i < 16; i++) L[i] = 0.0;
    for (int j = 0; j < 4; j++) {
        double d = S[j * 4 + j];
        for (int k = 0; k < j; k++) d -= L[j * 4 + k] * L[j * 4 + k];
        d = sqrt(d);
        L[j * 4 + j] = d;
        for (int i = j + 1; i < 4; i++) {
            double s = S[i * 4 + j];
            for (int k = 0; k < j; k++) s -= L[i * 4 + k] * L[j * 4 + k];
            L[i * 4 + j] = s / d;
        }
    }
    // gain[row] = solve(S, (P H^T)[row])  for the 8 rows of P[:, :4]
    double G[32];
    for (int row = 0; row < 8; row++) {
        double y[4];
        for (int i = 0; i < 4; i++) {
            double s = P[row * 8 + i];
            for (int k = 0; k < i; k++) s -= L[i * 4 + k] * y[k];
            y[i] = s / L[i * 4 + i];
        }
        for (int i = 3; i >= 0; i--) {
            double s = y[i];
            for (int k = i + 1; k < 4; k++) s -= L[k * 4 + i] * G[row * 4 + k];
            G[row * 4 + i] = s / L[i * 4 + i];
        }
    }
    double innov[4];
    for (int k = 0; k < 4; k++) innov[k] = (double)z32[k] - m[k];
    for (int j = 0; j < 8; j++) {
        double s = 0.0;
        for (int k = 0; k < 4; k++) s += innov[k] * G[j * 4 + k];
        m[j] = m[j] + s;
    }
    // P -= (G S) G^T
    double GS[32];
    for (int i = 0; i < 8; i++)
        for (int j = 0; j < 4; j++) {
            double s = 0.0;
            for (int k = 0; k < 4; k++) s += G[i * 4 + k] * S[k * 4 + j];
            GS[i * 4 + j] = s;
        }
    for (int i = 0; i < 8; i++)
        for (int j = 0; j < 8; j++) {
            double s = 0.0;
            for (int k = 0; k < 4; k++) s += GS[i * 4 + k] * G[j * 4 + k];
            P[i * 8 + j] = P[i * 8 + j] - s;
        }
}

BT_HD void bt_xyah32(const float* tlwh, float* z) {         // STrack.tlwh_to_xyah on a float32 row
    float x = tlwh[0], y = tlwh[1], w = tlwh[2], h = tlwh[3];
    float hw = w / 2.0f, hh = h / 2.0f;
    x = x + hw;
    y = y + hh;
    w = w / h;
    z[0] = x; z[1] = y; z[2] = w; z[3] = h;
}

// ------------------------------------------------------------------------------------------------ assignment
// scipy.optimize.linear_sum_assignment (rectangular_lsap.cpp), nr <= nc, C row-major nr x nc.  Lanes share the scan
// over the remaining columns; the winner of a scan is the one the serial loop would end with: lowest reduced cost; among
// equals the LAST unassigned column in scan order if there is one, else the FIRST.
struct BtCand { double s; int it; int un; };

BT_HD bool bt_cand_better(const BtCand& a, const BtCand& b) {      // a beats b
    if (b.it < 0) return a.it >= 0;
    if (a.it < 0) return false;
    if (a.s != b.s) return a.s < b.s;
    if (a.un != b.un) return a.un > b.un;
    return a.un ? a.it > b.it : a.it < b.it;
}

BT_HD int bt_lsap(int nr, int nc, const double* C, BtWork* w) {
    const int lane = BT_LANE;
    for (int j = lane; j < nc; j += BT_NL) { w->v[j] = 0.0; w->row4col[j] = -1; w->path[j] = -1; }
    for (int i = lane; i < nr; i += BT_NL) { w->u[i] = 0.0; w->col4row[i] = -1; }
    BT_SYNC();
    for (int cur = 0; cur < nr; cur++) {
        for (int j = lane; j < nc; j += BT_NL) { w->remaining[j] = nc - j - 1; w->SC[j] = 0; w->spc[j] = INFINITY; }
        for (int i = lane; i < nr; i += BT_NL) w->SR[i] = 0;
        BT_SYNC();
        int num_remaining = nc, sink = -1, i = cur;
        double minVal = 0.0;
        while (sink == -1) {
            if (lane == 0) w->SR[i] = 1;
            const double ui = w->u[i];
            BtCand best; best.s = INFINITY; best.it = -1; best.un = 0;
            for (int it = lane; it < num_remaining; it += BT_NL) {
                const int j = w->remaining[it];
                const double r = minVal + C[i * nc + j] - ui - w->v[j];
                double s = w->spc[j];
                if (r < s) { w->path[j] = i; w->spc[j] = r; s = r; }
                BtCand cand; cand.s = s; cand.it = it; cand.un = w->row4col[j] == -1;
                // serial rule: replace if s < lowest || (s == lowest && unassigned)
                if (best.it < 0 || cand.s < best.s || (cand.s == best.s && cand.un)) best = cand;
            }
#if defined(__CUDA_ARCH__)
            for (int off = 16; off > 0; off >>= 1) {
                BtCand o;
                o.s = __shfl_xor_sync(0xffffffffu, best.s, off);
                o.it = __shfl_xor_sync(0xffffffffu, best.it, off);
                o.un = __shfl_xor_sync(0xffffffffu, best.un, off);
                if (bt_cand_better(o, best)) best = o;
            }
#endif
            if (best.it < 0 || best.s == INFINITY) return -1;            // infeasible (cannot happen with finite costs)
            minVal = best.s;
            const int index = best.it;
            const int j = w->remaining[index];
            const int r4c = w->row4col[j];
            const int last = w->remaining[num_remaining - 1];
            BT_SYNC();
            if (lane == 0) { w->SC[j] = 1; w->remaining[index] = last; }
            num_remaining--;
            if (r4c == -1) sink = j; else i = r4c;
            BT_SYNC();
        }
        if (lane == 0) w->u[cur] += minVal;
        for (int k = lane; k < nr; k += BT_NL)
            if (w->SR[k] && k != cur) w->u[k] += minVal - w->spc[w->col4row[k]];
        for (int j = lane; j < nc; j += BT_NL)
            if (w->SC[j]) w->v[j] -= minVal - w->spc[j];
        BT_SYNC();
        if (lane == 0) {
            int j = sink;
            while (true) {
                const int k = w->path[j];
                w->row4col[j] = k;
                const int t = w->col4row[k];
                w->col4row[k] = j;
                j = t;
                if (k == cur) break;
            }
        }
        BT_SYNC();
    }
    return 0;
}

// matching.linear_assignment: clamp costs above thresh to thresh + 1e-4, solve, keep matches with cost <= thresh.
// Results in w->ma/mb (row order), w->ua, w->ub (ascending).
BT_HD void bt_assign(BtWork* w, int na, int nb, double* cost, double* costT, double thresh) {
    const int lane = BT_LANE;
    if (na == 0 || nb == 0) {
        if (lane == 0) {
            w->n_match = 0; w->n_ua = na; w->n_ub = nb;
            for (int i = 0; i < na; i++) w->ua[i] = (short)i;
            for (int j = 0; j < nb; j++) w->ub[j] = (short)j;
        }
        BT_SYNC();
        return;
    }
    const double clamp = thresh + 1e-4;
    const bool transpose = nb < na;
    for (int p = lane; p < na * nb; p += BT_NL) {
        double x = cost[p];
        if (x > thresh) x = clamp;
        cost[p] = x;
        if (transpose) { const int i = p / nb, j = p - i * nb; costT[j * na + i] = x; }
    }
    BT_SYNC();
    const int rc = transpose ? bt_lsap(nb, na, costT, w) : bt_lsap(na, nb, cost, w);
    if (lane == 0) {
        if (rc != 0) w->fail = 1;
        int nm = 0, nua = 0, nub = 0;
        // original row i is matched to: col4row[i] (plain) / row4col[i] (transposed problem: its columns are our rows)
        for (int j = 0; j < nb; j++) w->SC[j] = 0;
        for (int i = 0; i < na; i++) {
            const int j = rc != 0 ? -1 : (transpose ? w->row4col[i] : w->col4row[i]);
            if (j >= 0 && cost[i * nb + j] <= thresh) { w->ma[nm] = (short)i; w->mb[nm] = (short)j; nm++; w->SC[j] = 1; }
            else w->ua[nua++] = (short)i;
        }
        for (int j = 0; j < nb; j++) if (!w->SC[j]) w->ub[nub++] = (short)j;
        w->n_match = nm; w->n_ua = nua; w->n_ub = nub;
    }
    BT_SYNC();
}

// ------------------------------------------------------------------------------------------------ bookkeeping (lane 0)
BT_HD int bt_alloc_slot(BtClip* c) {
    for (int t = 0; t < BT_T; t++) if (!c->in_use[t]) { c->in_use[t] = 1; return t; }
    return -1;
}

BT_HD void bt_initiate(BtClip* c, const BtParams& p, int t, const float* tlwh) {       // STrack.activate
    float z[4];
    bt_xyah32(tlwh, z);
    const float h = z[3];
    const float k_pos = (float)(2 * BT_W_POS), k_vel = (float)(10 * BT_W_VEL);          // python float * np.float32 -> float32
    const float a = k_pos * h, b = k_vel * h;
    const double std[8] = {(double)a, (double)a, 1e-2, (double)a, (double)b, (double)b, 1e-5, (double)b};
    double* m = c->mean[t];
    double* P = c->cov[t];
    for (int i = 0; i < 4; i++) { m[i] = (double)z[i]; m[i + 4] = 0.0; }
    for (int i = 0; i < 64; i++) P[i] = 0.0;
    for (int i = 0; i < 8; i++) P[i * 8 + i] = std[i] * std[i];
    c->state[t] = BT_TRACKED;
    c->len[t] = 0;
    c->activated[t] = c->frame_id == 1;
    c->ext[t] = -1;
    if (p.min_consec == 1) c->ext[t] = c->next_external++;
    c->frame[t] = c->start[t] = c->frame_id;
}

// STrack.update / re_activate bookkeeping after the Kalman correction, in match order (ids are issued in that order)
BT_HD void bt_hit(BtClip* c, const BtParams& p, int t, bool was_tracked) {
    c->state[t] = BT_TRACKED;
    c->frame[t] = c->frame_id;
    if (!was_tracked) { c->len[t] = 0; return; }
    c->len[t] += 1;
    if (c->len[t] == p.min_consec) {
        c->activated[t] = 1;
        if (c->ext[t] == -1) c->ext[t] = c->next_external++;
    }
}

BT_HD bool bt_contains(const short* a, int n, int t) {
    for (int i = 0; i < n; i++) if (a[i] == t) return true;
    return false;
}

// ------------------------------------------------------------------------------------------------ one frame
// Detections of the frame: rows [0, n_rows) of xyxy / conf / cls (K2a's output layout for one image).  Writes the tracked
// detections (the rows supervision's `detections[tracker_id != -1]` keeps, in detection order) to out_row / out_tid and
// returns their number, or -1 if a capacity was exceeded.
BT_HD int bt_update(BtClip* c, BtWork* w, const BtParams& p, const float* xyxy, const float* conf, const int32_t* cls,
                    int n_rows, double* cost, double* costT, int32_t* out_row, int32_t* out_tid) {
    const int lane = BT_LANE;
    // ---- detections: the mask of main.py:189-193, then the high / low score split of update_with_tensors
    if (lane == 0) {
        w->fail = 0;
        int n0 = 0, nh = 0, nl = 0;
        for (int r = 0; r < n_rows; r++) {
            const int k = cls ? cls[r] : 0;
            const bool ok = conf[r] > p.min_conf && k >= 0 && k < 32 && ((p.class_mask >> k) & 1u);
            if (!ok) continue;
            if (n0 >= BT_D) { w->fail = 1; break; }
            w->row0[n0] = (short)r;
            const float s = conf[r];
            w->score[n0] = s;
            if (s > p.act_thr) w->hi[nh++] = (short)n0;
            else if (s > 0.1f && s < p.act_thr) w->lo[nl++] = (short)n0;
            n0++;
        }
        w->n0 = n0; w->n_hi = nh; w->n_lo = nl;
        c->frame_id += 1;
    }
    BT_SYNC();
    const int n0 = w->n0, n_hi = w->n_hi, n_lo = w->n_lo;
    for (int i = lane; i < n0; i += BT_NL) {
        const float* b = xyxy + 4 * (int)w->row0[i];
        const float x1 = b[0], y1 = b[1], x2 = b[2], y2 = b[3];
        w->xyxy[i][0] = x1; w->xyxy[i][1] = y1; w->xyxy[i][2] = x2; w->xyxy[i][3] = y2;
        const float bw = x2 - x1, bh = y2 - y1;                      // tlbr_to_tlwh (float32)
        w->tlwh[i][0] = x1; w->tlwh[i][1] = y1; w->tlwh[i][2] = bw; w->tlwh[i][3] = bh;
        w->box[i][0] = x1; w->box[i][1] = y1; w->box[i][2] = bw + x1; w->box[i][3] = bh + y1;   // STrack.tlbr (float32)
    }
    // ---- track lists
    if (lane == 0) {
        int nu = 0, np = 0;
        for (int i = 0; i < c->n_tracked; i++) { const int t = c->tracked[i]; if (!c->activated[t]) w->unc[nu++] = (short)t; }
        for (int i = 0; i < c->n_tracked; i++) { const int t = c->tracked[i]; if (c->activated[t]) w->pool[np++] = (short)t; }
        const int n_conf = np;
        for (int i = 0; i < c->n_lost; i++) {
            const int t = c->lost[i];
            if (bt_contains(w->pool, n_conf, t)) continue;
            if (np >= BT_T) { w->fail = 1; break; }
            w->pool[np++] = (short)t;
        }
        w->n_unc = nu; w->n_pool = np;
        w->n_act = w->n_refind = w->n_newlost = w->n_newrem = 0;
    }
    BT_SYNC();
    const int n_pool = w->n_pool, n_unc = w->n_unc;
    for (int i = lane; i < n_pool; i += BT_NL) bt_predict(c, w->pool[i]);
    BT_SYNC();

    // ---- first association: confirmed + lost tracks vs high-score detections, IoU fused with the score
    for (int i = lane; i < n_pool; i += BT_NL) bt_tlbr(c->mean[w->pool[i]], w->abox[i]);
    for (int j = lane; j < n_hi; j += BT_NL) {
        const int d = w->hi[j];
        for (int k = 0; k < 4; k++) w->bbox[j][k] = (double)w->box[d][k];
        w->bscore[j] = (double)w->score[d];
    }
    BT_SYNC();
    bt_cost_matrix(w, n_pool, n_hi, 2, true, cost);
    bt_assign(w, n_pool, n_hi, cost, costT, p.match_thr);
    int nm = w->n_match;
    for (int k = lane; k < nm; k += BT_NL) {
        const int t = w->pool[w->ma[k]], d = w->hi[w->mb[k]];
        float z[4];
        bt_xyah32(w->tlwh[d], z);
        w->was[k] = c->state[t] == BT_TRACKED;
        bt_correct(c, t, z);
    }
    BT_SYNC();
    if (lane == 0) {
        for (int k = 0; k < nm; k++) bt_hit(c, p, w->pool[w->ma[k]], w->was[k]);
        for (int k = 0; k < nm; k++) if (w->was[k]) w->act[w->n_act++] = w->pool[w->ma[k]];
        for (int k = 0; k < nm; k++) if (!w->was[k]) w->refind[w->n_refind++] = w->pool[w->ma[k]];
        // second association operands: still-tracked leftovers of the pool; remaining high-score detections for the third
        int nr = 0;
        for (int k = 0; k < w->n_ua; k++) { const int t = w->pool[w->ua[k]]; if (c->state[t] == BT_TRACKED) w->rest[nr++] = (short)t; }
        w->n_rest = nr;
        for (int k = 0; k < w->n_ub; k++) w->rem[k] = w->ub[k];        // indices into hi[]
        w->n_rem = w->n_ub;
    }
    BT_SYNC();

    // ---- second association: leftovers vs low-score detections
    const int n_rest = w->n_rest;
    for (int i = lane; i < n_rest; i += BT_NL) bt_tlbr(c->mean[w->rest[i]], w->abox[i]);
    for (int j = lane; j < n_lo; j += BT_NL) {
        const int d = w->lo[j];
        for (int k = 0; k < 4; k++) w->bbox[j][k] = (double)w->box[d][k];
    }
    BT_SYNC();
    bt_cost_matrix(w, n_rest, n_lo, 2, false, cost);
    bt_assign(w, n_rest, n_lo, cost, costT, 0.5);
    nm = w->n_match;
    for (int k = lane; k < nm; k += BT_NL) {
        float z[4];
        bt_xyah32(w->tlwh[w->lo[w->mb[k]]], z);
        bt_correct(c, w->rest[w->ma[k]], z);
    }
    BT_SYNC();
    if (lane == 0) {
        for (int k = 0; k < nm; k++) { const int t = w->rest[w->ma[k]]; bt_hit(c, p, t, true); w->act[w->n_act++] = (short)t; }
        for (int k = 0; k < w->n_ua; k++) {
            const int t = w->rest[w->ua[k]];
            if (c->state[t] != BT_LOST) { c->state[t] = BT_LOST; w->newlost[w->n_newlost++] = (short)t; }
        }
    }
    BT_SYNC();

    // ---- third association: unconfirmed tracks vs the remaining high-score detections
    const int n_rem = w->n_rem;
    for (int i = lane; i < n_unc; i += BT_NL) bt_tlbr(c->mean[w->unc[i]], w->abox[i]);
    for (int j = lane; j < n_rem; j += BT_NL) {
        const int d = w->hi[w->rem[j]];
        for (int k = 0; k < 4; k++) w->bbox[j][k] = (double)w->box[d][k];
        w->bscore[j] = (double)w->score[d];
    }
    BT_SYNC();
    bt_cost_matrix(w, n_unc, n_rem, 2, true, cost);
    bt_assign(w, n_unc, n_rem, cost, costT, 0.7);
    nm = w->n_match;
    for (int k = lane; k < nm; k += BT_NL) {
        float z[4];
        bt_xyah32(w->tlwh[w->hi[w->rem[w->mb[k]]]], z);
        bt_correct(c, w->unc[w->ma[k]], z);
    }
    BT_SYNC();
    if (lane == 0) {
        for (int k = 0; k < nm; k++) { const int t = w->unc[w->ma[k]]; bt_hit(c, p, t, true); w->act[w->n_act++] = (short)t; }
        for (int k = 0; k < w->n_ua; k++) { const int t = w->unc[w->ua[k]]; c->state[t] = BT_REMOVED; w->newrem[w->n_newrem++] = (short)t; }
        // births
        for (int k = 0; k < w->n_ub; k++) {
            const int d = w->hi[w->rem[w->ub[k]]];
            if (w->score[d] < (float)p.det_thr) continue;
            const int t = bt_alloc_slot(c);
            if (t < 0 || w->n_act >= BT_T) { w->fail = 1; break; }
            bt_initiate(c, p, t, w->tlwh[d]);
            w->act[w->n_act++] = (short)t;
        }
        // lost for too long
        for (int i = 0; i < c->n_lost; i++) {
            const int t = c->lost[i];
            if (c->frame_id - c->frame[t] > p.max_time_lost) { c->state[t] = BT_REMOVED; if (w->n_newrem < BT_T) w->newrem[w->n_newrem++] = (short)t; }
        }
        // list updates (joint_tracks / sub_tracks of update_with_tensors); w->pool is free to be reused as a temporary
        int nt = 0;
        for (int i = 0; i < c->n_tracked; i++) { const int t = c->tracked[i]; if (c->state[t] == BT_TRACKED) c->tracked[nt++] = (short)t; }
        for (int pass = 0; pass < 2; pass++) {
            const short* src = pass == 0 ? w->act : w->refind;
            const int ns = pass == 0 ? w->n_act : w->n_refind;
            for (int i = 0; i < ns; i++) {
                if (bt_contains(c->tracked, nt, src[i])) continue;
                if (nt >= BT_T) { w->fail = 1; break; }
                c->tracked[nt++] = src[i];
            }
        }
        c->n_tracked = nt;
        int nl = 0;
        for (int i = 0; i < c->n_lost; i++) { const int t = c->lost[i]; if (!bt_contains(c->tracked, nt, t)) w->pool[nl++] = (short)t; }
        for (int i = 0; i < w->n_newlost; i++) { if (nl >= BT_T) { w->fail = 1; break; } w->pool[nl++] = w->newlost[i]; }
        int nl2 = 0;
        for (int i = 0; i < nl; i++) if (!bt_contains(c->removed, c->n_removed, w->pool[i])) c->lost[nl2++] = w->pool[i];
        c->n_lost = nl2;
        for (int i = 0; i < w->n_newrem; i++) c->removed[i] = w->newrem[i];
        c->n_removed = w->n_newrem;
    }
    BT_SYNC();

    // ---- duplicates between tracked and lost (IoU distance < 0.15): keep the longer-lived one
    {
        const int nt = c->n_tracked, nl = c->n_lost;
        if (nt > 0 && nl > 0) {
            for (int i = lane; i < nt; i += BT_NL) bt_tlbr(c->mean[c->tracked[i]], w->abox[i]);
            for (int j = lane; j < nl; j += BT_NL) bt_tlbr(c->mean[c->lost[j]], w->bbox[j]);
            BT_SYNC();
            bt_cost_matrix(w, nt, nl, 0, false, cost);
            if (lane == 0) {
                for (int i = 0; i < nt; i++) w->SR[i] = 0;
                for (int j = 0; j < nl; j++) w->SC[j] = 0;
                for (int i = 0; i < nt; i++)
                    for (int j = 0; j < nl; j++)
                        if (cost[i * nl + j] < 0.15) {
                            const int ta = c->tracked[i], tb = c->lost[j];
                            if (c->frame[ta] - c->start[ta] > c->frame[tb] - c->start[tb]) w->SC[j] = 1; else w->SR[i] = 1;
                        }
                int a = 0, b = 0;
                for (int i = 0; i < nt; i++) if (!w->SR[i]) c->tracked[a++] = c->tracked[i];
                for (int j = 0; j < nl; j++) if (!w->SC[j]) c->lost[b++] = c->lost[j];
                c->n_tracked = a; c->n_lost = b;
            }
            BT_SYNC();
        }
    }

    // ---- slots: in use = tracked + lost + removed-this-frame (the latter still filter `lost` next frame)
    for (int t = lane; t < BT_T; t += BT_NL) c->in_use[t] = 0;
    BT_SYNC();
    if (lane == 0) {
        for (int i = 0; i < c->n_tracked; i++) c->in_use[c->tracked[i]] = 1;
        for (int i = 0; i < c->n_lost; i++) c->in_use[c->lost[i]] = 1;
        for (int i = 0; i < c->n_removed; i++) c->in_use[c->removed[i]] = 1;
        int no = 0;
        for (int i = 0; i < c->n_tracked; i++) { const int t = c->tracked[i]; if (c->activated[t]) w->out[no++] = (short)t; }
        w->n_out = no;
    }
    BT_SYNC();

    // ---- update_with_detections: put tracker ids on the detections (all filtered detections vs activated tracked tracks)
    const int n_out = w->n_out;
    int kept = 0;
    if (n_out > 0 && n0 > 0) {
        for (int i = lane; i < n0; i += BT_NL)
            for (int k = 0; k < 4; k++) w->abox[i][k] = (double)w->xyxy[i][k];
        for (int j = lane; j < n_out; j += BT_NL) bt_tlbr(c->mean[w->out[j]], w->bbox[j]);
        BT_SYNC();
        bt_cost_matrix(w, n0, n_out, 1, false, cost);
        bt_assign(w, n0, n_out, cost, costT, 0.5);
        if (lane == 0) {
            for (int i = 0; i < n0; i++) w->ids[i] = -1;
            for (int k = 0; k < w->n_match; k++) w->ids[w->ma[k]] = c->ext[w->out[w->mb[k]]];
            int n = 0;
            for (int i = 0; i < n0; i++)
                if (w->ids[i] != -1) { out_row[n] = w->row0[i]; out_tid[n] = w->ids[i]; n++; }
            w->n_match = n;
        }
        BT_SYNC();
        kept = w->n_match;
    }
    BT_SYNC();
    if (w->fail) { if (lane == 0) c->overflow = 1; BT_SYNC(); return -1; }
    return kept;
}

BT_HD void bt_reset(BtClip* c) {       // all lanes
    const int lane = BT_LANE;
    if (lane == 0) { c->frame_id = 0; c->next_external = 1; c->n_tracked = c->n_lost = c->n_removed = 0; c->overflow = 0; c->next_seq = 0; }
    for (int t = lane; t < BT_T; t += BT_NL) { c->in_use[t] = 0; c->state[t] = BT_NEW; c->activated[t] = 0; c->ext[t] = -1; }
    BT_SYNC();
}
