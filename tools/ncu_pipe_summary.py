"""Tensor-pipe / memory headline metrics per launch from an `ncu --set full` report.
usage: python tools/ncu_pipe_summary.py report.ncu-rep [more.ncu-rep ...]   (needs ncu on PATH; no GPU)"""
import csv
import io
import subprocess
import sys

WANT = ("gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
        "sm__pipe_tensor_cycles_active_realtime.avg.pct_of_peak_sustained_elapsed",
        "sm__pipe_tensor_subpipe_hmma_cycles_active_realtime.avg", "sm__inst_executed_pipe_tensor.sum", "sm__inst_executed_pipe_uniform.sum",
        "sm__inst_executed_pipe_tc.sum", "sm__inst_executed_pipe_tmem.sum", "sm__cycles_elapsed.max", "sm__warps_active.avg.pct_of_peak_sustained_active",
        "l1tex__data_pipe_lsu_wavefronts.avg.pct_of_peak_sustained_elapsed", "lts__t_bytes.sum", "launch__registers_per_thread",
        "launch__grid_size", "launch__block_size", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
        "smsp__inst_executed.sum")


def main():
    for path in sys.argv[1:]:
        raw = subprocess.run(["ncu", "-i", path, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
        rows = list(csv.reader(io.StringIO(raw)))
        hdr, units = rows[0], rows[1]
        names = [h.split(".TriageCompute.")[-1] if ".Triage" in h else h for h in hdr]
        print("## %s\n" % path)
        for r in rows[2:]:
            d = dict(zip(names, r))
            print("```\n  %s" % d.get("Kernel Name", "?")[:110])
            for i, n in enumerate(names):
                if n in WANT or n in ("sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active", "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed",
                                      "sm__pipe_tensor_subpipe_hmma_cycles_active.avg.pct_of_peak_sustained_active",
                                      "sm__inst_executed_pipe_tensor_subpipe_hmma.avg.pct_of_peak_sustained_active"):
                    print("  %-86s %s %s" % (n, r[i], units[i]))
            print("```")


if __name__ == "__main__":
    main()
