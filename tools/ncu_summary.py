"""Turn an `ncu --set full` report into the small JSON summaries kept under profiles/ (run here, no GPU needed).

    python tools/ncu_summary.py gemm  gpurun_out/X_gemm.ncu-rep  profiles/rNN_ncu_gemm_MODE.json  [patches] [passes-note]
    python tools/ncu_summary.py front gpurun_out/X_front.ncu-rep profiles/rNN_ncu_front.json      [patches]
"""
import collections, csv, io, json, subprocess, sys

MUL = {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}
TIME = {"ns": 1e-3, "us": 1.0, "usecond": 1.0, "ms": 1e3, "msecond": 1e3, "s": 1e6, "second": 1e6, "nsecond": 1e-3}


def page(rep, which, extra=()):
    out = subprocess.run(["ncu", "-i", rep, "--page", which, "--csv", *extra], capture_output=True, text=True).stdout
    return list(csv.reader(io.StringIO(out)))


def raw_rows(rep):
    rows = page(rep, "raw")
    hdr, units = rows[0], rows[1]
    ix = {h: i for i, h in enumerate(hdr)}

    def get(r, name, kind=None):
        v = float(r[ix[name]].replace(",", ""))
        u = units[ix[name]]
        if kind == "bytes":
            return v * MUL.get(u, 1)
        if kind == "us":
            return v * TIME.get(u, 1)
        return v
    return hdr, rows[2:], get, ix


def gemm(rep, out, patches=4000, note=""):
    names = ["conv2", "conv3_1", "conv3_2", "conv4_1", "conv4_2", "fc1", "fc2", "fc3"]
    mmac = [113.246, 113.246, 226.492, 113.246, 226.492, 50.332, 16.777, 0.524]      # SURVEY B.2, per patch
    hdr, rows, get, ix = raw_rows(rep)
    layers = []
    for i, r in enumerate(rows[:8]):
        dur = get(r, "gpu__time_duration.sum", "us")
        rd, wr = get(r, "dram__bytes_read.sum", "bytes"), get(r, "dram__bytes_write.sum", "bytes")
        fl = 2 * mmac[i] * 1e6 * patches
        layers.append({"layer": names[i], "kernel": r[ix["Kernel Name"]][:64], "duration_us": dur, "dram_read_bytes": rd,
                       "dram_write_bytes": wr, "algorithmic_tflops": fl / dur / 1e6,
                       "tensor_pipe_active_pct_of_elapsed": get(r, "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed")
                       if "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed" in ix else None,
                       "registers": get(r, "launch__registers_per_thread"), "grid": get(r, "launch__grid_size")})
    tot_d = sum(l["duration_us"] for l in layers)
    tot_b = sum(l["dram_read_bytes"] + l["dram_write_bytes"] for l in layers)
    json.dump({"source": f"ncu --set full --clock-control none --import-source on -k regex:fadb_gemm_tc_kernel -s 24 -c 8 "
                         f"python tools/prof_front.py {patches // 10}  ({note})",
               "patches": patches, "layers": layers, "sum_duration_us": tot_d, "dram_bytes_8_launches": tot_b,
               "dram_bytes_per_patch": tot_b / patches,
               "algorithmic_tflops_all": sum(2 * m * 1e6 * patches for m in mmac) / tot_d / 1e6}, open(out, "w"), indent=1)
    for l in layers:
        print(l["layer"], round(l["duration_us"], 1), round(l["algorithmic_tflops"]), l["tensor_pipe_active_pct_of_elapsed"])
    print("sum", tot_d, "us; DRAM bytes per patch", tot_b / patches)


def front(rep, out, patches=2040, title=""):
    hdr, rows, get, ix = raw_rows(rep)
    r = rows[0]
    keys = ["gpu__time_duration.sum", "sm__cycles_elapsed.avg", "dram__bytes_read.sum", "dram__bytes_write.sum",
            "l1tex__data_pipe_lsu_wavefronts.avg.pct_of_peak_sustained_elapsed", "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum",
            "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum.pct_of_peak_sustained_elapsed",
            "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active",
            "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active",
            "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active",
            "smsp__issue_active.avg.pct_of_peak_sustained_active", "sm__warps_active.avg.pct_of_peak_sustained_active",
            "launch__registers_per_thread", "launch__shared_mem_per_block_dynamic", "launch__grid_size", "smsp__inst_executed.sum"]
    m = {k: float(r[ix[k]].replace(",", "")) for k in keys if k in ix}
    units = page(rep, "raw")[1]
    stalls = {h.replace("smsp__average_warps_issue_stalled_", "").replace("_per_issue_active.ratio", ""): float(r[i].replace(",", ""))
              for h, i in ix.items() if h.startswith("smsp__average_warps_issue_stalled_") and h.endswith("_per_issue_active.ratio")}
    sass = page(rep, "source", ("--print-source", "sass"))
    h2 = sass[1]
    jx = {h: i for i, h in enumerate(h2)}
    samp, ex, wf = collections.Counter(), collections.Counter(), collections.Counter()
    for row in sass[2:]:
        if len(row) < len(h2):
            continue
        t = row[jx["Source"]].split()
        if not t:
            continue
        op = (t[1] if t[0].startswith("@") else t[0]).split(".")[0]
        samp[op] += float(row[jx["# Samples"]] or 0)
        ex[op] += float(row[jx["Instructions Executed"]] or 0)
        wf[op] += float(row[jx["L1 Wavefronts Shared"]] or 0)
    tot = sum(samp.values()) or 1
    dur = get(r, "gpu__time_duration.sum", "us")
    dram = get(r, "dram__bytes_read.sum", "bytes") + get(r, "dram__bytes_write.sum", "bytes")
    json.dump({"kernel": "fadb_vggish_front_conv1_tc_kernel " + title,
               "command": f"ncu --set full --clock-control none --import-source on -k regex:fadb_vggish_front -s 3 -c 1 python tools/prof_front.py {patches // 10}",
               "patches": patches, "duration_us": dur, "metrics": m,
               "stall_cycles_per_issued_instruction": dict(sorted(stalls.items(), key=lambda kv: -kv[1])),
               "top_opcodes_by_stall_samples": [{"opcode": k, "share": v / tot, "warp_instructions": ex[k], "shared_wavefronts": wf[k]}
                                                for k, v in samp.most_common(12)],
               "per_patch": {"duration_us_per_patch_per_sm": dur * 148 / patches,
                             "shared_wavefronts": m.get("l1tex__data_pipe_lsu_wavefronts_mem_shared.sum", 0) / patches,
                             "warp_instructions": m.get("smsp__inst_executed.sum", 0) / patches,
                             "dram_bytes": dram / patches, "algorithmic_bytes_survey_8d": 88576.0, "actual_io_bytes": 260608.0}},
              open(out, "w"), indent=1)
    print("duration", dur, "us;", m.get("l1tex__data_pipe_lsu_wavefronts_mem_shared.sum", 0) / patches, "shared wavefronts / patch;",
          m.get("smsp__inst_executed.sum", 0) / patches, "warp instr / patch")


if __name__ == "__main__":
    kind, rep, out = sys.argv[1:4]
    n = int(sys.argv[4]) if len(sys.argv) > 4 else (4000 if kind == "gemm" else 2040)
    (gemm if kind == "gemm" else front)(rep, out, n, sys.argv[5] if len(sys.argv) > 5 else "")
