"""Turn an ncu report into the short text summary committed under profiles/.

    python profiles/summarize_ncu.py gpurun_out/prof.ncu-rep > profiles/rNN_name.txt

Prints the headline metrics of every captured launch (duration, DRAM bytes, throughput percentages, IPC,
occupancy) and, from the SASS source page, instructions / stall-sample shares of the code between barriers.
"""
import csv
import io
import subprocess
import sys

METRICS = [
    "gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
    "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
    "l1tex__throughput.avg.pct_of_peak_sustained_elapsed", "lts__throughput.avg.pct_of_peak_sustained_elapsed",
    "sm__warps_active.avg.pct_of_peak_sustained_active", "sm__inst_executed.avg.per_cycle_elapsed",
    "smsp__inst_executed.sum", "sm__cycles_elapsed.avg", "sm__cycles_active.avg", "launch__registers_per_thread",
    "launch__shared_mem_per_block_dynamic", "launch__occupancy_limit_shared_mem", "launch__occupancy_limit_registers",
    "launch__grid_size", "launch__block_size", "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum",
    "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum", "smsp__warps_eligible.avg.per_cycle_active",
]


def page(rep, name):
    out = subprocess.run(["ncu", "-i", rep, "--page", name, "--csv"], capture_output=True, text=True).stdout
    return list(csv.reader(io.StringIO(out)))


def main(rep):
    rows = page(rep, "raw")
    hdr, units, data = rows[0], rows[1], rows[2:]
    name_col = hdr.index("Kernel Name")
    print(f"# ncu summary of {rep}")
    for li, d in enumerate(data):
        print(f"\n## launch {li}: {d[name_col][:110]}")
        for m in METRICS:
            if m in hdr:
                i = hdr.index(m)
                print(f"{m:72s} {d[i]:>18s} {units[i]}")
    src = page(rep, "source")
    if len(src) < 3:
        return
    h = src[1]
    ci = {k: i for i, k in enumerate(h)}
    stalls = [k for k in h if k.startswith("stall_") and "Not Issued" not in k]
    segs, cur = [], None

    def fresh():
        return dict(inst=0, samples=0, sass=0, **{k: 0 for k in stalls})

    cur = fresh()
    for r in src[2:]:
        if len(r) < len(h) or r[0] in ("Address", "Kernel Name"):
            if r and r[0] == "Kernel Name":
                break
            continue
        try:
            inst, smp = int(r[ci["Instructions Executed"]]), int(r[ci["# Samples"]])
        except ValueError:
            continue
        cur["inst"] += inst
        cur["samples"] += smp
        cur["sass"] += 1
        for k in stalls:
            try:
                cur[k] += int(r[ci[k]])
            except ValueError:
                pass
        if "BAR.SYNC" in r[ci["Source"]] or "EXIT" in r[ci["Source"]]:
            segs.append(cur)
            cur = fresh()
    segs.append(cur)
    ti, ts = sum(s["inst"] for s in segs) or 1, sum(s["samples"] for s in segs) or 1
    print("\n## first launch, SASS between barriers (segments with >= 1 % of the samples)")
    print("seg  sass  warp-inst%  samples%  top stall reasons (% of the segment's stall samples)")
    for i, s in enumerate(segs):
        if s["samples"] < 0.01 * ts:
            continue
        tot = sum(s[k] for k in stalls) or 1
        top = sorted(stalls, key=lambda k: -s[k])[:4]
        print(f"{i:3d} {s['sass']:5d} {100 * s['inst'] / ti:9.1f} {100 * s['samples'] / ts:9.1f}   "
              + ", ".join(f"{k[6:]} {100 * s[k] / tot:.0f}" for k in top))


if __name__ == "__main__":
    main(sys.argv[1])
