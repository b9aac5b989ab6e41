"""Turn the round-2 ncu captures (gpurun_out/r02p_*, written by tools/refresh_profiles_r02.sh on the GPU box) into the
summaries committed under profiles/: key metrics per kernel (JSON), executed instructions per opcode and per source line,
stall reasons, the launch list of the bench command, and profiles/r02_summary.json (DRAM traffic per launch of the
benchmarked kernel, read by bench.py for roofline.traffic).

    python tools/summarise_profiles_r02.py        (build machine; needs cuobjdump / nvdisasm for the line tables)
"""
import csv
import json
import os
import shutil
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
SRC, DST = os.path.join(ROOT, "gpurun_out"), os.path.join(ROOT, "profiles")
LIB = os.path.join(ROOT, "gif-synthesis-with-discrete-diffusion_b200", "csrc", "libd3pm_b200.so")

CAPTURES = {
    # tag: (mangled-name fragment, divisor = warp-rows of the launch, what)
    "stream_k4096_on": ("step_stream_kernelILi4ELi8ELb1ELb0", 16 * 4096 * 4, "config 2: 16 videos x 4096 tokens, K = 4096, guidance 2 (4 warps per row)"),
    "stream_k4096_off": ("step_stream_kernelILi4ELi16ELb0ELb0", 32 * 4096 * 2, "32 videos, K = 4096, guidance off (2 warps per row)"),
    "stream_k2048_on": ("step_stream_kernelILi2ELi8ELb1ELb0", 32 * 4096 * 2, "32 videos, K = 2048, guidance 2 (2 warps per row)"),
    "stream_k1024_on": ("step_stream_kernelILi1ELi8ELb1ELb0", 64 * 4096 * 1, "64 videos, K = 1024, guidance 2 (1 warp per row)"),
    "head": ("head_step_kernelILi64ELb1ELb0", 16 * 4096 * 4096 / 32, "fused head + step, 16 x 4096 tokens x 4096 classes (divisor: warp-level (row, class) pairs)"),
    "train": ("train_stream_kernelILi4ELb1", 16 * 1024 * 4, "losses + gradient, 16 x 1024 tokens x 4096 codes (4 warps per row)"),
}
METRICS = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "dram__throughput.avg.pct_of_peak_sustained_elapsed",
           "smsp__issue_active.avg.pct_of_peak_sustained_active", "smsp__inst_executed.sum", "launch__registers_per_thread",
           "launch__block_size", "launch__grid_size", "smsp__warps_active.avg.per_cycle_active", "lts__t_sector_hit_rate.pct",
           "sm__inst_executed_pipe_xu.sum", "sm__inst_executed_pipe_tensor.sum", "smsp__inst_executed_pipe_uniform.sum"]


def raw_metrics(path):
    rows = list(csv.reader(open(path)))
    hdr, unit, val = rows[0], rows[1], rows[-1]
    out = {"kernel": val[hdr.index("Kernel Name")]}
    for m in METRICS:
        if m in hdr:
            i = hdr.index(m)
            out[m] = {"value": val[i], "unit": unit[i]}
    return out


def to_bytes(entry):
    v, u = float(entry["value"].replace(",", "")), entry["unit"].lower()
    return v * {"byte": 1, "kbyte": 1e3, "mbyte": 1e6, "gbyte": 1e9}[u]


def run(cmd, out):
    with open(out, "w") as f:
        subprocess.run(cmd, stdout=f, stderr=subprocess.DEVNULL, check=False)


def main():
    tmp = "/tmp/r02_sass"
    os.makedirs(tmp, exist_ok=True)
    subprocess.run(["cuobjdump", "-xelf", "all", LIB], cwd=tmp, stdout=subprocess.DEVNULL, check=True)
    cubin = [f for f in os.listdir(tmp) if f.endswith(".cubin")][0]
    run(["nvdisasm", "-g", "-c", os.path.join(tmp, cubin)], os.path.join(tmp, "all.sass"))
    summary = {}
    for tag, (frag, div, what) in CAPTURES.items():
        raw = os.path.join(SRC, f"r02p_{tag}_raw.csv")
        sass = os.path.join(SRC, f"r02p_{tag}_sass.csv")
        if not os.path.isfile(raw):
            print("missing", raw)
            continue
        m = raw_metrics(raw)
        m["what"], m["plain_run"] = what, open(os.path.join(SRC, f"r02p_{tag}_plain.log")).read().strip().splitlines()[-1][:400]
        m["dram_bytes_per_launch"] = to_bytes(m["dram__bytes_read.sum"]) + to_bytes(m["dram__bytes_write.sum"])
        summary[tag] = m
        shutil.copyfile(raw, os.path.join(DST, f"r02_{tag}_ncu_raw.csv"))
        run([sys.executable, os.path.join(ROOT, "tools", "sass_hist.py"), sass, str(div)], os.path.join(DST, f"r02_{tag}_sass_hist.txt"))
        run([sys.executable, os.path.join(ROOT, "tools", "sass_stalls.py"), sass], os.path.join(DST, f"r02_{tag}_stalls.txt"))
        run([sys.executable, os.path.join(ROOT, "tools", "sass_lines.py"), sass, os.path.join(tmp, "all.sass"), frag, str(div), "0.004"],
            os.path.join(DST, f"r02_{tag}_lines.txt"))
    json.dump(summary, open(os.path.join(DST, "r02_kernels.json"), "w"), indent=1)
    k = summary["stream_k4096_on"]
    json.dump({"traffic_bytes_per_launch": k["dram_bytes_per_launch"],
               "source": "ncu --set full --clock-control none, profiles/r02_stream_k4096_on_ncu_raw.csv (step_stream_kernel<4, 8, true, false>, config 2)",
               "algorithmic_bytes_per_launch": 16 * 4096 * 32784}, open(os.path.join(DST, "r02_summary.json"), "w"), indent=1)
    # launch list of `python bench.py --steps 5 --warmup 3`
    src = os.path.join(SRC, "r02p_bench_launches.csv")
    if os.path.isfile(src):
        lines = [l for l in open(src) if not l.startswith("==")]
        open(os.path.join(DST, "r02_bench_launches.csv"), "w").writelines(lines)
    shutil.copyfile(os.path.join(SRC, "r02p_bench_plain.log"), os.path.join(DST, "r02_bench_line_steps5.json"))
    for tag, m in summary.items():
        print(tag, m["gpu__time_duration.sum"]["value"], m["gpu__time_duration.sum"]["unit"], "dram %.4g B" % m["dram_bytes_per_launch"],
              "issue", m.get("smsp__issue_active.avg.pct_of_peak_sustained_active", {}).get("value"))


if __name__ == "__main__":
    main()
