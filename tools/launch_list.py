"""Turn an ncu launch list (`--metrics gpu__time_duration.sum --csv`) of ONE bench step into the markdown
summary committed under profiles/: libhvb kernels vs library kernels, launches, total time and share.

    python tools/launch_list.py gpurun_out/launches_r01d.csv --title "..." > profiles/r01d_launch_list.md
"""
import argparse
import collections
import csv

HVB = ("letterbox_kernel", "decode_nms_kernel", "crops_from_boxes_kernel", "color_features_kernel", "mnv3_prep",
       "scale_transform_kernel", "standardize", "gram_tcgen05", "affinity_from_gram", "d2_f64", "iou_cost_kernel",
       "merge_nms_kernel", "gather_tiles_kernel", "bias_act_kernel", "concat_nhwc_kernel", "stem_conv_kernel", "sppf_pool_concat_kernel", "cvt_hsv_lab",
       "pointwise_conv_kernel", "jersey_color_stats_kernel", "bytetrack_kernel", "select_crops_kernel")


def short(name):
    name = name.replace("void ", "").replace("<unnamed>::", "").replace("at::native::", "at::")
    p = name.find("(")
    name = name[:p] if p > 0 else name
    return name if len(name) <= 90 else name[:87] + "..."


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("csv")
    ap.add_argument("--title", default="Launch list of one timed bench step")
    ap.add_argument("--command", default="")
    ap.add_argument("--note", default="")
    a = ap.parse_args()
    rows = list(csv.reader(open(a.csv)))
    hdr = next(i for i, r in enumerate(rows) if r and r[0] == "ID")
    h = rows[hdr]
    mi, vi, ui = h.index("Metric Name"), h.index("Metric Value"), h.index("Metric Unit")
    scale = {"ns": 1e-3, "nsecond": 1e-3, "us": 1.0, "usecond": 1.0, "ms": 1e3, "msecond": 1e3,
             "byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}
    launches = collections.OrderedDict()                      # ID -> [name, us, dram bytes]
    for r in rows[hdr + 1:]:
        if len(r) <= vi:
            continue
        rec = launches.setdefault(r[0], [short(r[4]), 0.0, 0.0])
        v = float(r[vi].replace(",", "")) * scale.get(r[ui], 1.0)
        if r[mi] == "gpu__time_duration.sum":
            rec[1] = v
        elif r[mi].startswith("dram__bytes"):
            rec[2] += v
    data = list(launches.values())
    agg = collections.OrderedDict()
    total = 0.0
    for k, us, nbytes in data:
        total += us
        n, t, b = agg.get(k, (0, 0.0, 0.0))
        agg[k] = (n + 1, t + us, b + nbytes)
    have_bytes = any(b for _, _, b in agg.values())
    ours = {k: v for k, v in agg.items() if any(h in k for h in HVB)}
    lib = {k: v for k, v in agg.items() if k not in ours}
    print("# %s\n" % a.title)
    if a.command:
        print("Command (after the same command exited 0 without ncu):\n\n```\n%s\n```\n" % a.command)
    if a.note:
        print(a.note + "\n")
    print("Times are cold-cache and serialised by ncu: compare SHARES, not absolutes.  %d launches, %.2f ms summed.\n" % (len(data), total / 1e3))
    print("## libhvb kernels (hand-written, sm_100a)\n")
    print("| kernel | launches | total us | share of step |%s\n|---|---|---|---|%s" % (" DRAM MB (read+write) | DRAM GB/s |" if have_bytes else "", "---|---|" if have_bytes else ""))
    so = 0.0
    for k, (n, t, b) in sorted(ours.items(), key=lambda kv: -kv[1][1]):
        extra = " %.1f | %.0f |" % (b / 1e6, b / t / 1e3) if have_bytes else ""
        print("| `%s` | %d | %.1f | %.2f %% |%s" % (k, n, t, 100 * t / total, extra))
        so += t
    print("| **all libhvb** | %d | %.1f | **%.1f %%** |\n" % (sum(v[0] for v in ours.values()), so, 100 * so / total))
    print("## Library kernels (PyTorch / cuDNN / CUTLASS convolutions and the few remaining torch ops)\n")
    print("| kernel | launches | total us | share |%s\n|---|---|---|---|%s" % (" DRAM MB | DRAM GB/s |" if have_bytes else "", "---|---|" if have_bytes else ""))
    for k, (n, t, b) in sorted(lib.items(), key=lambda kv: -kv[1][1])[:25]:
        extra = " %.1f | %.0f |" % (b / 1e6, b / t / 1e3) if have_bytes else ""
        print("| `%s` | %d | %.1f | %.1f %% |%s" % (k, n, t, 100 * t / total, extra))
    sl = sum(v[1] for v in lib.values())
    print("| **all library** | %d | %.1f | **%.1f %%** |" % (sum(v[0] for v in lib.values()), sl, 100 * sl / total))


if __name__ == "__main__":
    main()
