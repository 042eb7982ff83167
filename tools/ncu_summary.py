"""Summarise an .ncu-rep (run where ncu is installed, no GPU needed): key raw metrics per captured launch,
stall-reason shares, and the source lines that execute the most instructions.
    python tools/ncu_summary.py gpurun_out/prof.ncu-rep [--updates N] [--top 30] > profiles/xyz.md"""
import argparse, collections, csv, io, re, subprocess, sys

KEYS = ["gpu__time_duration.sum", "launch__grid_size", "launch__block_size", "launch__registers_per_thread",
        "launch__occupancy_limit_registers", "sm__warps_active.avg.pct_of_peak_sustained_active",
        "smsp__inst_executed.sum", "smsp__issue_active.avg.pct_of_peak_sustained_active",
        "sm__throughput.avg.pct_of_peak_sustained_elapsed", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "dram__throughput.avg.pct_of_peak_sustained_elapsed", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
        "lts__t_sector_hit_rate.pct", "lts__t_sectors.sum", "lts__throughput.avg.pct_of_peak_sustained_elapsed",
        "lts__t_requests_srcunit_tex_op_read.sum", "lts__t_sectors_srcunit_tex_op_read.sum",
        "lts__t_sectors_srcunit_tex_op_read_lookup_hit.sum", "lts__t_sectors_srcunit_tex_op_read_lookup_miss.sum",
        "lts__t_sectors_srcunit_tex_op_red.sum", "lts__t_sectors_srcunit_ltcfabric.sum",
        "l1tex__t_sectors_pipe_lsu_mem_global_op_ld.sum", "l1tex__t_sectors_pipe_lsu_mem_global_op_red.sum",
        "l1tex__t_sector_hit_rate.pct", "l1tex__m_l1tex2xbar_req_cycles_active.avg.pct_of_peak_sustained_elapsed"]


def ncu(args):
    return subprocess.run(["ncu"] + args, capture_output=True, text=True).stdout


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("rep")
    ap.add_argument("--updates", type=float, default=0, help="applied updates per launch (adds per-update columns)")
    ap.add_argument("--top", type=int, default=30)
    a = ap.parse_args()
    raw = list(csv.reader(io.StringIO(ncu(["-i", a.rep, "--page", "raw", "--csv"]))))
    hdr, units, launches = raw[0], raw[1], raw[2:]
    print(f"# ncu summary of `{a.rep}`\n")
    for li, vals in enumerate(launches):
        print(f"## launch {li}: {vals[hdr.index('Kernel Name')][:90]}\n")
        print("| metric | value | unit |" + (" per update |" if a.updates else ""))
        print("|---|---:|---|" + ("---:|" if a.updates else ""))
        for k in KEYS:
            if k in hdr:
                i = hdr.index(k)
                v = vals[i]
                extra = ""
                if a.updates:
                    try:
                        f = float(v.replace(",", ""))
                        scale = {"Tbyte": 1e12, "Gbyte": 1e9, "Mbyte": 1e6, "Kbyte": 1e3}.get(units[i], 1.0)
                        extra = f" {f * scale / a.updates:.2f} |" if (".sum" in k and "time" not in k) else " |"
                    except ValueError:
                        extra = " |"
                print(f"| {k} | {v} | {units[i]} |{extra}")
        print()
    src = list(csv.reader(io.StringIO(ncu(["-i", a.rep, "--page", "source", "--csv", "--print-source", "cuda,sass"]))))
    out, stalls, cur, h, fn, nfn = [], collections.Counter(), None, None, None, 0
    for r in src:
        if not r:
            continue
        if r[0] == "File Path":
            cur = r[1].split("/")[-1]; continue
        if r[0] == "Function Name":
            if r[1] != fn:
                fn = r[1]; nfn += 1
            continue
        if r[0] == "Line No":
            h = r; continue
        if nfn > 1:
            break            # first captured launch only
        if r[0] != "" and h:
            try:
                out.append((float(r[h.index("Instructions Executed")]), float(r[h.index("# Samples")]), cur, r[0], r[1].strip()[:100]))
                for k, name in enumerate(h):
                    if name.startswith("stall_") and "Not Issued" not in name:
                        stalls[name] += float(r[k] or 0)
            except ValueError:
                pass
    tot, ts = sum(o[0] for o in out), sum(o[1] for o in out)
    if tot:
        print(f"## source hot spots (first launch): {tot:.4g} warp instructions" +
              (f" = {tot / a.updates * 32:.0f} per 32 updates" if a.updates else "") + "\n")
        ss = sum(stalls.values())
        print("stall reasons: " + ", ".join(f"{k[6:]} {100 * v / ss:.1f}%" for k, v in stalls.most_common(8)) + "\n")
        print("| inst % | samples % | where | source |\n|---:|---:|---|---|")
        for o in sorted(out, reverse=True)[:a.top]:
            print(f"| {100 * o[0] / tot:.1f} | {100 * o[1] / ts:.1f} | {o[2]}:{o[3]} | `{o[4].replace('|', '/')}` |")


if __name__ == "__main__":
    main()
