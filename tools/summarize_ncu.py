#!/usr/bin/env python
"""Turn an `ncu --set full --import-source on` report into the markdown / json summaries kept under
profiles/ (run where ncu is installed; no GPU needed):

    python tools/summarize_ncu.py gpurun_out/prof.ncu-rep profiles/r02_kernels --traffic-key gen_kernel \
        --real-per-launch 1000 --traffic-out profiles/r02_traffic.json

Every kernel in the report gets a metric table, its SASS opcode mix and warp-state samples; the traffic json
(dram bytes per launch of the --traffic-key kernel) is stamped with the hash of the kernel sources it was captured from
(bench.py compares it with the build it runs and marks `roofline.traffic` stale otherwise).
"""
import collections
import csv
import io
import json
import subprocess
import sys

KEYS = [
    "gpu__time_duration.sum", "launch__grid_size", "launch__block_size", "launch__registers_per_thread",
    "launch__shared_mem_per_block_dynamic", "launch__occupancy_limit_registers", "launch__occupancy_limit_shared_mem",
    "sm__warps_active.avg.pct_of_peak_sustained_active", "smsp__issue_active.avg.pct_of_peak_sustained_active",
    "smsp__warps_eligible.avg.per_cycle_active", "smsp__inst_executed.sum", "sm__cycles_elapsed.max",
    "dram__bytes_read.sum", "dram__bytes_write.sum", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
    "lts__t_bytes.sum", "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum",
    "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active", "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active",
    "sm__pipe_shared_cycles_active.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_tensor_subpipe_imma.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_tensor_subpipe_dmma.avg.pct_of_peak_sustained_active",
    "lts__throughput.avg.pct_of_peak_sustained_elapsed", "l1tex__throughput.avg.pct_of_peak_sustained_elapsed",
    "sm__throughput.avg.pct_of_peak_sustained_elapsed",
]


def ncu_csv(rep, page, kernel=None):
    cmd = ["ncu", "-i", rep, "--page", page, "--csv"] + (["--kernel-name", "regex:" + kernel] if kernel else [])
    out = subprocess.run(cmd, capture_output=True, text=True).stdout
    return list(csv.reader(io.StringIO(out)))


def source_hash():
    import hashlib
    import os
    d = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "pta_replicator_b200", "csrc")
    h = hashlib.sha256()
    for f in sorted(os.listdir(d)):
        if f in ("ptar_generate.cuh", "ptar_rng.cuh"):      # the device code of the captured kernel (gen_kernel)
            with open(os.path.join(d, f), "rb") as fh:
                h.update(fh.read())
    return h.hexdigest()[:16]


def to_bytes(v, unit):
    f = float(v)
    return f * {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "Tbyte": 1e12}.get(unit, 1)


def main():
    rep, stem = sys.argv[1], sys.argv[2]
    tkey = sys.argv[sys.argv.index("--traffic-key") + 1] if "--traffic-key" in sys.argv else None
    rpl = int(sys.argv[sys.argv.index("--real-per-launch") + 1]) if "--real-per-launch" in sys.argv else None
    raw = ncu_csv(rep, "raw")
    hdr, units = raw[0], raw[1]
    md = [f"# ncu summary of `{rep.split('/')[-1]}`  (--set full --clock-control none)\n"]
    traffic = {}
    for r in raw[2:]:
        d = {h: (r[i], units[i]) for i, h in enumerate(hdr)}
        name = d["Kernel Name"][0]
        md.append(f"## `{name[:110]}`\n")
        md.append("| metric | value | unit |\n|---|---|---|")
        for k in KEYS:
            if k in d and d[k][0] != "":
                md.append(f"| {k} | {d[k][0]} | {d[k][1]} |")
        for k, (v, u) in d.items():
            if "issue_stalled" in k and k.endswith("per_issue_active.ratio"):
                try:
                    if float(v) >= 0.3:
                        md.append(f"| {k} | {v} | warps |")
                except ValueError:
                    pass
        rd = to_bytes(*d["dram__bytes_read.sum"])
        wr = to_bytes(*d["dram__bytes_write.sum"])
        md.append(f"\nDRAM traffic per launch: {rd + wr:.4g} B (read {rd:.4g} + write {wr:.4g})\n")
        if tkey and tkey in name:
            traffic[f"{tkey}_dram_bytes_per_launch"] = rd + wr
            dur, du = d["gpu__time_duration.sum"]
            traffic[f"{tkey}_duration_us_under_ncu"] = float(dur) * {"ns": 1e-3, "us": 1.0, "ms": 1e3, "s": 1e6}.get(du, 1.0)
            traffic[f"{tkey}_grid"] = d["launch__grid_size"][0]
            if rpl:
                traffic[f"{tkey}_realizations_per_launch"] = rpl
    seen = set()
    for r in raw[2:]:
        name = dict(zip(hdr, r))["Kernel Name"]
        short = name.split("(")[0].split("<")[0].split("::")[-1].split()[-1]
        if short in seen:
            continue
        seen.add(short)
        src = ncu_csv(rep, "source", short)
        hi = [i for i, r2 in enumerate(src) if r2 and r2[0] == "Address"]
        if not hi:
            continue
        h = src[hi[0]]
        data = src[hi[0] + 1:(hi[1] - 1 if len(hi) > 1 else len(src))]
        ix = {n: i for i, n in enumerate(h)}
        op, ops = collections.Counter(), collections.Counter()
        tot_i = tot_s = 0
        for r2 in data:
            if len(r2) <= ix["# Samples"]:
                continue
            toks = r2[ix["Source"]].strip().split()
            if not toks:
                continue
            o = (toks[1] if toks[0].startswith("@") and len(toks) > 1 else toks[0]).split(".")[0]
            try:
                ni, ns = int(r2[ix["Instructions Executed"]] or 0), int(r2[ix["# Samples"]] or 0)
            except ValueError:
                continue
            op[o] += ni; ops[o] += ns; tot_i += ni; tot_s += ns
        md.append(f"### `{short}`: SASS opcode mix (warp instructions executed, share of stall samples)\n")
        md.append("| opcode | executed | % instr | % samples |\n|---|---|---|---|")
        for o, c in op.most_common(22):
            md.append(f"| {o} | {c} | {100 * c / max(tot_i, 1):.1f} | {100 * ops[o] / max(tot_s, 1):.1f} |")
        st = [n for n in h if n.startswith("stall_") and "Not Issued" not in n]
        tot = {}
        for n in st:
            t = 0
            for r2 in data:
                if len(r2) > ix[n]:
                    try:
                        t += int(r2[ix[n]] or 0)
                    except ValueError:
                        pass
            tot[n] = t
        md.append("\nwarp-state samples: " + ", ".join(f"{k[6:]} {100 * v / max(tot_s, 1):.1f}%" for k, v in sorted(tot.items(), key=lambda kv: -kv[1])[:10]))
        marks = [o for o in op if o in ("UBLKCP", "UTMALDG", "DMMA", "LDGSTS", "SYNCS", "UTCHMMA", "UTCIMMA", "LDTM", "UTCBAR")]
        md.append("\nBlackwell/Hopper-class instructions executed: " + (", ".join(f"{m} x{op[m]}" for m in marks) or "none") + "\n")
    open(stem + ".md", "w").write("\n".join(md) + "\n")
    if traffic:
        traffic["source_hash"] = source_hash()
        tout = sys.argv[sys.argv.index("--traffic-out") + 1] if "--traffic-out" in sys.argv else stem + "_traffic.json"
        open(tout, "w").write(json.dumps(traffic, indent=1) + "\n")
    print("wrote", stem + ".md")


if __name__ == "__main__":
    main()
