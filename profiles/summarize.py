#!/usr/bin/env python
"""Turns an `ncu --set full` report (.ncu-rep, read on the CPU box with `ncu -i`) into the small text summary that is
committed under profiles/ (the .ncu-rep itself stays in gpurun_out/, which is scratch).

    python profiles/summarize.py gpurun_out/prof_X.ncu-rep profiles/r01_X.md [--source]

Per launch: duration, DRAM bytes (the bench's roofline.traffic), L1/L2 hit rates, achieved L1/L2/DRAM throughput, occupancy,
registers, warp execution efficiency (threads per instruction), issue utilisation and the warp-stall breakdown.
--source adds a per-code-region table (share of samples / instructions / no-instruction stalls) from the source page.
"""
import csv
import io
import json
import subprocess
import sys

KEYS = [
    ("gpu__time_duration.sum", "duration"),
    ("launch__grid_size", "grid"),
    ("launch__block_size", "block"),
    ("launch__registers_per_thread", "registers/thread"),
    ("launch__shared_mem_per_block_dynamic", "dynamic smem/block"),
    ("launch__occupancy_limit_registers", "occupancy limit (registers), blocks/SM"),
    ("launch__occupancy_limit_shared_mem", "occupancy limit (smem), blocks/SM"),
    ("sm__warps_active.avg.pct_of_peak_sustained_active", "achieved occupancy % of max warps"),
    ("smsp__inst_executed.sum", "warp instructions executed"),
    ("smsp__thread_inst_executed_per_inst_executed.ratio", "warp execution efficiency (active threads / instruction, of 32)"),
    ("smsp__issue_active.avg.pct_of_peak_sustained_active", "issue slots busy %"),
    ("smsp__warps_eligible.avg.per_cycle_active", "eligible warps / cycle / scheduler"),
    ("sm__throughput.avg.pct_of_peak_sustained_elapsed", "SM throughput %"),
    ("l1tex__t_sector_hit_rate.pct", "L1 hit rate %"),
    ("lts__t_sector_hit_rate.pct", "L2 hit rate %"),
    ("l1tex__t_bytes.sum", "L1 bytes"),
    ("lts__t_bytes.sum", "L2 bytes"),
    ("lts__t_sectors_op_read.sum", "L2 read sectors"),
    ("lts__t_sectors_op_write.sum", "L2 write sectors"),
    ("dram__bytes_read.sum", "DRAM read"),
    ("dram__bytes_write.sum", "DRAM write"),
    ("gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "DRAM throughput % of peak"),
    ("lts__throughput.avg.pct_of_peak_sustained_elapsed", "L2 throughput % of peak"),
    ("l1tex__throughput.avg.pct_of_peak_sustained_elapsed", "L1 throughput % of peak"),
    ("smsp__sass_average_data_bytes_per_sector_mem_global_op_ld.ratio", "bytes used per 32-B sector, global loads"),
    ("l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "shared-memory bank conflicts"),
    ("smsp__sass_inst_executed_op_shared_ld.sum", "shared loads (warp inst)"),
    ("smsp__sass_inst_executed_op_shared_st.sum", "shared stores (warp inst)"),
    ("smsp__sass_inst_executed_op_local_ld.sum", "local loads (warp inst)"),
    ("smsp__sass_inst_executed_op_local_st.sum", "local stores (warp inst)"),
    ("smsp__average_warp_latency_per_inst_issued.ratio", "warp latency per instruction issued (cycles)"),
]
STALL_PREFIX = "smsp__average_warps_issue_stalled_"


def raw(path):
    out = subprocess.run(["ncu", "-i", path, "--page", "raw", "--csv"], capture_output=True, text=True, check=True).stdout
    rows = list(csv.reader(io.StringIO(out)))
    return rows[0], rows[1], rows[2:]


def source(path):
    out = subprocess.run(["ncu", "-i", path, "--page", "source", "--csv"], capture_output=True, text=True, check=True).stdout
    kernels, cur = [], None
    for r in csv.reader(io.StringIO(out)):
        if r and r[0] == "Kernel Name":
            cur = {"name": r[1], "hdr": None, "rows": []}
            kernels.append(cur)
        elif cur is not None and cur["hdr"] is None:
            cur["hdr"] = r
        elif cur is not None:
            cur["rows"].append(r)
    return kernels


def main():
    rep, out = sys.argv[1], sys.argv[2]
    want_source = "--source" in sys.argv
    hdr, units, launches = raw(rep)
    ix = {k: i for i, k in enumerate(hdr)}
    lines = [f"# ncu summary of `{rep}`", "", "Captured with `ncu --set full --clock-control none --import-source on` under gpurun (B200, sm_100a); "
             "read on the CPU box with `ncu -i … --page raw --csv`. Times under ncu are serialised/cold-cache: never a bench value.", ""]
    summary = []
    for r in launches:
        name = r[ix["Kernel Name"]]
        lines += [f"## {name}", "", "| metric | value | unit |", "|---|---:|---|"]
        d = {"kernel": name}
        for k, label in KEYS:
            if k in ix and r[ix[k]] != "":
                lines.append(f"| {label} (`{k}`) | {r[ix[k]]} | {units[ix[k]]} |")
                d[k] = r[ix[k]]
        stalls = sorted(((float(r[i]), k[len(STALL_PREFIX):].replace("_per_issue_active.ratio", "")) for k, i in ix.items()
                         if k.startswith(STALL_PREFIX) and k.endswith("_per_issue_active.ratio") and r[i] not in ("", "0")), reverse=True)
        lines += ["", "Warp stall reasons (warps stalled per issue-active cycle, largest first): " +
                  ", ".join(f"{n} {v:.2f}" for v, n in stalls[:9]), ""]
        d["stalls"] = {n: v for v, n in stalls}
        try:
            rd, wr = float(r[ix["dram__bytes_read.sum"]]), float(r[ix["dram__bytes_write.sum"]])
            scale = {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}
            d["dram_bytes"] = rd * scale[units[ix["dram__bytes_read.sum"]]] + wr * scale[units[ix["dram__bytes_write.sum"]]]
            lines += [f"DRAM traffic (read + write) per launch: **{d['dram_bytes'] / 1e6:.1f} MB**", ""]
        except Exception:
            pass
        summary.append(d)
    if want_source:
        seen = set()
        for k in source(rep):
            if k["name"] in seen:
                continue
            seen.add(k["name"])
            h = {n: i for i, n in enumerate(k["hdr"])}
            rows = k["rows"]
            tot_s = sum(int(r[h["# Samples"]]) for r in rows) or 1
            tot_i = sum(int(r[h["Instructions Executed"]]) for r in rows) or 1
            lines += [f"## source page, {k['name']}: {len(rows)} SASS instructions, {tot_i} warp instructions, {tot_s} samples", "",
                      "| SASS rows | % samples | % warp inst | no_inst % of region samples | long_sb % | wait % | threads/inst | first instruction |",
                      "|---|---:|---:|---:|---:|---:|---:|---|"]
            B = 100
            for b in range(0, len(rows), B):
                seg = rows[b:b + B]
                s = sum(int(r[h["# Samples"]]) for r in seg)
                ie = sum(int(r[h["Instructions Executed"]]) for r in seg)
                th = sum(int(r[h["Thread Instructions Executed"]]) for r in seg)
                f = lambda c: 100.0 * sum(int(r[h[c]]) for r in seg) / max(s, 1)
                if s * 200 < tot_s and ie * 200 < tot_i:
                    continue
                lines.append(f"| {b}-{b + len(seg) - 1} | {100 * s / tot_s:.1f} | {100 * ie / tot_i:.1f} | {f('stall_no_inst'):.0f} | {f('stall_long_sb'):.0f} | "
                             f"{f('stall_wait'):.0f} | {th / max(ie, 1):.1f} | `{seg[0][1].strip()[:48]}` |")
            lines.append("")
    with open(out, "w") as f:
        f.write("\n".join(lines))
    with open(out.rsplit(".", 1)[0] + ".json", "w") as f:
        json.dump(summary, f, indent=1)
    print(f"wrote {out}")


if __name__ == "__main__":
    main()
