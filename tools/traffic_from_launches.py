#!/usr/bin/env python
"""profiles/launches_r2_<workload>.csv (ncu launch list with dram bytes) -> profiles/traffic_r2.json:
measured DRAM bytes per launch of every kernel class of bench.py's `kernels` object."""
import collections, csv, json, os, re, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
CLASS = [("spmv_pattern_multi", "spmv_aux"), ("spmv_sell_multi", "spmv_aux"), ("spmv_", "spmv"), ("reduce_partials", "spmv"),
         ("gram_kernel", "mdot"), ("mdotm_kernel", "mdot"), ("mdot_kernel", "mdot"), ("mdot_reg_kernel", "mdot"),
         ("lincomb_kernel", "lincomb"), ("lincomb2_kernel", "lincomb"), ("lincomb2n_kernel", "lincomb"), ("orth_mid_kernel", "orthmid"),
         ("scale_kernel", "scale"), ("hess_kernel", "other"), ("pipe_init_kernel", "other")]
MODE = re.compile(r"kernel<\(?(?:int\))?\s*(\d)")
out = {}
for wl, nrows in (("lkdv", 10_000_050), ("swe", 10_002_828)):
    path = os.path.join(ROOT, "profiles", f"launches_r2_{wl}.csv")
    rows = [r for r in csv.reader(open(path)) if len(r) > 10]
    hdr, rows = rows[0], rows[1:]
    ik, im, iv, iid = (hdr.index(k) for k in ("Kernel Name", "Metric Name", "Metric Value", "ID"))
    per = collections.OrderedDict()
    for r in rows:
        per.setdefault(r[iid], {"k": r[ik]})[r[im]] = float(r[iv].replace(",", ""))
    agg = collections.defaultdict(lambda: dict(launches=0, ns=0.0, dram=0.0))
    # single-vector mode-0 SpMVs: the first one of the solve (Arnoldi step 0) is on the system matrix; launches
    # that read clearly less ran on a constraint matrix (a third of A's rows carry entries)
    ref_read = None
    for d in per.values():
        name = d["k"]
        cls = next((c for pat, c in CLASS if pat in name), None)
        if cls is None:
            continue
        dram = d.get("dram__bytes_read.sum", 0.0) + d.get("dram__bytes_write.sum", 0.0)
        single0 = cls == "spmv" and "dual" not in name and "reduce_partials" not in name and MODE.search(name) and MODE.search(name).group(1) == "0"
        if single0:
            rd = d.get("dram__bytes_read.sum", 0.0)
            if ref_read is None:
                ref_read = rd
            elif abs(rd - ref_read) > 0.2 * ref_read:
                cls = "spmv_aux"
        a = agg[cls]; a["launches"] += 0 if "reduce_partials" in name else 1; a["ns"] += d["gpu__time_duration.sum"]; a["dram"] += dram
    tot = sum(a["ns"] for a in agg.values())
    out[wl] = {c: {"launches": a["launches"], "traffic_bytes_per_launch": a["dram"] / a["launches"],
                   "ncu_us_per_launch": a["ns"] / a["launches"] * 1e-3, "share_of_kernel_time": a["ns"] / tot,
                   "dram_gbs_under_ncu": a["dram"] / a["ns"]} for c, a in agg.items()}
    out[wl]["_source"] = f"profiles/launches_r2_{wl}.csv (ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none; one solve)"
json.dump(out, open(os.path.join(ROOT, "profiles", "traffic_r2.json"), "w"), indent=1)
for wl in out:
    for c, v in out[wl].items():
        if c[0] != "_":
            print(wl, c, {k: round(x, 3) if isinstance(x, float) else x for k, x in v.items()})
