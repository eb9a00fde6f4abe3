"""Per-kernel summary of an `ncu --set full` report, in the form of the files under profiles/:
    python tools/ncu_summary.py gpurun_out/<name>.ncu-rep [--stalls] > profiles/<name>_summary.txt
--stalls adds, per kernel, where the warps wait (mbarrier waits by barrier address, fences, named barriers) from the SASS
page of the report (needs --import-source on at capture time)."""
import csv
import io
import subprocess
import sys

METRICS = [
    ("duration", "gpu__time_duration.sum"),
    ("SM cycles elapsed", "sm__cycles_elapsed.max"),
    ("average SM clock", "sm__cycles_elapsed.avg.per_second"),
    ("tensor pipe active % (of elapsed cycles)", "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed"),
    ("SM throughput %", "sm__throughput.avg.pct_of_peak_sustained_elapsed"),
    ("issue slots busy %", "smsp__issue_active.avg.pct_of_peak_sustained_active"),
    ("DRAM read", "dram__bytes_read.sum"),
    ("DRAM write", "dram__bytes_write.sum"),
    ("L2 -> SM read", "l1tex__m_xbar2l1tex_read_bytes.sum"),
    ("registers / thread", "launch__registers_per_thread"),
    ("grid", "launch__grid_size"),
    ("cluster", "launch__cluster_dim_x"),
    ("block", "launch__block_size"),
    ("dynamic smem / block", "launch__shared_mem_per_block_dynamic"),
    ("warp instructions executed", "smsp__inst_executed.sum"),
]


def page(rep, name, extra=()):
    out = subprocess.run(["ncu", "-i", rep, "--page", name, "--csv", *extra], capture_output=True, text=True).stdout
    return list(csv.reader(io.StringIO(out)))


def short(kname):
    return kname.replace("void ", "").replace("dmip::", "").replace("<unnamed>::", "").replace("(int)", "")


def main():
    rep = sys.argv[1]
    rows = page(rep, "raw")
    hdr, units = rows[0], rows[1]
    ik = hdr.index("Kernel Name")
    print(f"ncu -i {rep} --page raw  (one row per captured launch; numbers under the profiler are not bench values)")
    for r in rows[2:]:
        print(f"== {short(r[ik])}")
        for label, m in METRICS:
            if m in hdr:
                print(f"   {label:44s}{r[hdr.index(m)]:>22s} {units[hdr.index(m)]}")
        st = [(float(r[i]), h) for i, h in enumerate(hdr)
              if h.startswith("smsp__average_warps_issue_stalled") and h.endswith("per_issue_active.ratio")]
        top = sorted(st, reverse=True)[:5]
        print("   top stall reasons (warps per issue)         " + ", ".join(
            f"{h.replace('smsp__average_warps_issue_stalled_', '').replace('_per_issue_active.ratio', '')} {v:.2f}" for v, h in top))
    if "--stalls" in sys.argv:
        srows = page(rep, "source", ["--print-source", "sass"])
        # the page is a sequence of blocks: ["Kernel Name", name], header row, instruction rows
        i = 0
        while i < len(srows):
            if srows[i] and srows[i][0] == "Kernel Name":
                kname, h = srows[i][1], srows[i + 1]
                j = i + 2
                while j < len(srows) and not (srows[j] and srows[j][0] == "Kernel Name"):
                    j += 1
                body = srows[i + 2:j]
                i_s, i_src = h.index("# Samples"), h.index("Source")
                data = []
                for r in body:
                    try:
                        data.append((int(r[i_s]), r[i_src]))
                    except (ValueError, IndexError):
                        pass
                tot = sum(d[0] for d in data) or 1
                print(f"== where the warps of {short(kname)} wait ({tot} samples)")
                for n, (s, src) in enumerate(data):
                    nxt = data[n + 1][0] if n + 1 < len(data) else 0
                    if "TRYWAIT" in src and s + nxt > 0.003 * tot:
                        print(f"   {100 * (s + nxt) / tot:5.1f} %  mbarrier wait   {src.strip()[:80]}")
                    elif any(t in src for t in ("MEMBAR", "FENCE.VIEW", "ERRBAR", "BAR.SYNC")) and s + nxt > 0.003 * tot:
                        print(f"   {100 * (s + nxt) / tot:5.1f} %  fence / barrier {src.strip()[:80]}")
                i = j
            else:
                i += 1


if __name__ == "__main__":
    main()
