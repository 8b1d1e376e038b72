"""Summarise .ncu-rep captures (read here, without a GPU) into small text files for profiles/.
    python tools/ncu_summary.py gpurun_out/prof_x.ncu-rep [...]  > profiles/r01_x.txt
Per kernel launch: duration, grid, registers, DRAM bytes read/written + throughput, L2 throughput, tensor-pipe
utilisation, warps active, top stall reasons, and the hottest SASS instructions by stall samples."""
import csv
import io
import subprocess
import sys

KEYS = [
    ("gpu__time_duration.sum", "duration"),
    ("launch__grid_size", "grid"),
    ("launch__block_size", "block"),
    ("launch__registers_per_thread", "regs/thread"),
    ("launch__shared_mem_per_block_dynamic", "dyn smem/block"),
    ("launch__waves_per_multiprocessor", "waves/SM"),
    ("dram__bytes_read.sum", "dram read"),
    ("dram__bytes_write.sum", "dram write"),
    ("gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "dram % of peak"),
    ("lts__throughput.avg.pct_of_peak_sustained_elapsed", "L2 % of peak"),
    ("sm__throughput.avg.pct_of_peak_sustained_elapsed", "SM % of peak"),
    ("sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active", "tensor pipe % (active)"),
    ("sm__inst_executed_pipe_tensor.sum", "tensor instructions"),
    ("sm__warps_active.avg.pct_of_peak_sustained_active", "warps active %"),
    ("smsp__inst_executed.sum", "warp instructions"),
]


def run(args):
    return subprocess.run(["ncu"] + args, capture_output=True, text=True).stdout


def main():
    for rep in sys.argv[1:]:
        print(f"==== {rep}")
        raw = list(csv.reader(io.StringIO(run(["-i", rep, "--page", "raw", "--csv"]))))
        if len(raw) < 3:
            print("  (no data)")
            continue
        H, units = raw[0], raw[1]
        for r in raw[2:]:
            d = dict(zip(H, r))
            u = dict(zip(H, units))
            print(f"-- {d.get('Kernel Name', '?')[:110]}")
            for k, label in KEYS:
                if k in d:
                    print(f"   {label:26s} {d[k]} {u.get(k, '')}")
            stalls = []
            for h in H:
                if h.startswith("smsp__average_warps_issue_stalled_") and h.endswith("_per_issue_active.ratio"):
                    try:
                        stalls.append((float(d[h]), h[len("smsp__average_warps_issue_stalled_"):-len("_per_issue_active.ratio")]))
                    except ValueError:
                        pass
            stalls.sort(reverse=True)
            print("   stalls (warps per issue):  " + ", ".join(f"{n}={v:.2f}" for v, n in stalls[:6]))
        src = list(csv.reader(io.StringIO(run(["-i", rep, "--page", "source", "--csv"]))))
        # first kernel only
        if len(src) > 2 and "# Samples" in src[1]:
            Hs = src[1]
            ci, si = Hs.index("# Samples"), Hs.index("Source")
            rows = []
            for r in src[2:]:
                if r and r[0] == "Kernel Name":
                    break
                try:
                    rows.append((int(r[ci]), r[si].strip()[:100]))
                except (ValueError, IndexError):
                    pass
            tot = sum(x[0] for x in rows) or 1
            print(f"   hottest SASS by stall samples (first launch, {tot} samples):")
            for n, ins in sorted(rows, reverse=True)[:12]:
                print(f"     {100.0 * n / tot:5.1f}%  {ins}")


if __name__ == "__main__":
    main()
