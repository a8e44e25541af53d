"""Summarise an `ncu --set full` report (.ncu-rep) into a small tracked text table for profiles/.

  python tools/ncu_summary.py gpurun_out/prof.ncu-rep > profiles/r01_hpsi_full.md

Per captured launch: duration, DRAM bytes (read + write = `traffic`), DRAM / SM / L1 throughput %, pipe utilisation
(FP64 vector pipe and the tensor pipe, which executes DMMA), registers, occupancy, shared-memory
wavefronts and bank conflicts, and the top warp-stall reasons.
"""
import csv
import io
import subprocess
import sys


def raw(rep):
    out = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(out)))
    return rows[0], rows[1], rows[2:]


def f(x):
    try:
        return float(x.replace(",", ""))
    except Exception:
        return float("nan")


CLASS_OF = {"k_plane<0": "fft_plane", "k_plane_vloc": "fft_plane", "k_zpass_g2r": "fft_zpass", "k_zpass_r2g": "fft_zpass",
            "k_zgemm<1, 1": "gemm_project", "k_zgemm<0, 0": "gemm_expand", "k_shift_fused": "shift_fused",
            "k_shift_apply": "shift_fused", "k_shift_gemm": "shift_gemm", "k_plane_rho": "rho_plane"}


def traffic(rep, units_per_launch):
    """bench.py's profiles/traffic.json: measured DRAM bytes per unit (vector; RHS x outer iteration for the shifted
    update) and pipe utilisation of every kernel class, averaged over the captured launches."""
    import json
    hdr, units, data = raw(rep)
    ix = {h: i for i, h in enumerate(hdr)}
    acc = {}
    for d in data:
        name = d[ix["Kernel Name"]].replace("void ", "").replace("sgw::", "")
        cls = next((c for k, c in CLASS_OF.items() if name.startswith(k)), None)
        if cls is None:
            continue
        def byt(m):
            return f(d[ix[m]]) * {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}.get(units[ix[m]], 1.0)
        a = acc.setdefault(cls, {"n": 0, "bytes": 0.0, "fp64": 0.0, "tensor": 0.0, "dram": 0.0, "smem": 0.0})
        a["n"] += 1
        a["bytes"] += byt("dram__bytes_read.sum") + byt("dram__bytes_write.sum")
        a["fp64"] += f(d[ix["sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active"]])
        a["tensor"] += f(d[ix["TPC.TriageCompute.sm__pipe_tensor_cycles_active_realtime.avg.pct_of_peak_sustained_elapsed"]])
        a["dram"] += f(d[ix["gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed"]])
        if "l1tex__data_pipe_lsu_wavefronts.avg.pct_of_peak_sustained_elapsed" in ix:
            a["smem"] += f(d[ix["l1tex__data_pipe_lsu_wavefronts.avg.pct_of_peak_sustained_elapsed"]])
    out = {}
    for cls, a in acc.items():
        out[cls] = {"dram_bytes_per_unit": a["bytes"] / a["n"] / units_per_launch, "fp64_pipe_pct": a["fp64"] / a["n"],
                    "tensor_pipe_pct": a["tensor"] / a["n"], "dram_pct": a["dram"] / a["n"], "smem_pipe_pct": a["smem"] / a["n"],
                    "launches_captured": a["n"],
                    "source": rep.split("/")[-1], "units_per_launch": units_per_launch}
    print(json.dumps(out, indent=1).replace('NaN', 'null'))


def main():
    if sys.argv[1] == "--traffic":
        return traffic(sys.argv[3], float(sys.argv[2]))
    rep = sys.argv[1]
    hdr, units, data = raw(rep)
    ix = {h: i for i, h in enumerate(hdr)}

    def g(d, name):
        return f(d[ix[name]]) if name in ix else float("nan")

    def unit(name):
        return units[ix[name]] if name in ix else ""

    print(f"# ncu --set full summary of `{rep}`\n")
    print("| # | kernel | grid x block | time [us] | DRAM rd+wr [MB] | DRAM % | SM % | L1/smem % | FP64 pipe % | tensor (DMMA) pipe % | regs | "
          "occ % (theor.) | smem wavefronts | bank conflicts | top stalls |")
    print("|---|---|---|---|---|---|---|---|---|---|---|---|---|---|---|")
    for n, d in enumerate(data):
        name = d[ix["Kernel Name"]].split("(")[0].replace("void ", "")
        t = g(d, "gpu__time_duration.sum")
        tu = unit("gpu__time_duration.sum")
        t_us = t * {"ns": 1e-3, "us": 1.0, "ms": 1e3, "s": 1e6}.get(tu, 1e-3)
        def bytes_of(m):
            v, u = g(d, m), unit(m)
            return v * {"byte": 1e-6, "Kbyte": 1e-3, "Mbyte": 1.0, "Gbyte": 1e3}.get(u, 1e-6)
        dram = bytes_of("dram__bytes_read.sum") + bytes_of("dram__bytes_write.sum")
        stalls = [(f(d[i]), h.replace("smsp__pcsamp_warps_issue_stalled_", "")) for i, h in enumerate(hdr)
                  if "pcsamp_warps_issue_stalled" in h and "not_issued" not in h and d[i] not in ("", "n/a")]
        tot = sum(v for v, _ in stalls) or 1.0
        top = ", ".join(f"{h} {100 * v / tot:.0f}%" for v, h in sorted(stalls, reverse=True)[:3])
        print(f"| {n} | {name} | {int(g(d, 'launch__grid_size'))} x {int(g(d, 'launch__block_size'))} | {t_us:.1f} | {dram:.1f} | "
              f"{g(d, 'gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed'):.1f} | "
              f"{g(d, 'sm__throughput.avg.pct_of_peak_sustained_elapsed'):.1f} | "
              f"{g(d, 'l1tex__throughput.avg.pct_of_peak_sustained_active'):.1f} | "
              f"{g(d, 'sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active'):.1f} | "
              f"{g(d, 'TPC.TriageCompute.sm__pipe_tensor_cycles_active_realtime.avg.pct_of_peak_sustained_elapsed'):.1f} | "
              f"{int(g(d, 'launch__registers_per_thread'))} | "
              f"{g(d, 'sm__warps_active.avg.pct_of_peak_sustained_active'):.0f} ({g(d, 'sm__maximum_warps_per_active_cycle_pct'):.0f}) | "
              f"{g(d, 'l1tex__data_pipe_lsu_wavefronts_mem_shared.sum'):.3g} | "
              f"{g(d, 'l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum'):.3g} | {top} |")


if __name__ == "__main__":
    main()
