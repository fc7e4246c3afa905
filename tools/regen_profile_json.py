"""Regenerate profiles/tensor_pipe.json and profiles/traffic.json from .ncu-rep captures (no hand-maintained numbers):

    python tools/regen_profile_json.py gpurun_out            # directory holding ncu_<tag>.ncu-rep

MAP names which capture backs which entry; an entry whose capture is missing keeps its old value and is marked stale.
tensor_pipe.json: sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active of ONE cold launch (not a bench value);
traffic.json: dram__bytes_read.sum + dram__bytes_write.sum of that launch, keyed by bench.py's kernel names."""
import csv
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
# entry -> (capture tag, bench kernel name for traffic.json or None)
MAP = {
    "gru_decoder1_recurrence_H128_cfg2": ("r2_gru_dec1", "gru_decoder1_recurrence"),
    "gru_decoder2_step_H128_cfg2": ("r2_gru_dec2", "gru_decoder2_step"),
    "gru_decoder1_recurrence_H256_cfg3": ("r2_gru_dec1_cfg3", None),
    "social_fc_gemm_H128_cfg2": ("r2_social_fc", "social_fc_gemm"),
    "cvae_deconv3_gemm": ("r2_deconv3", "cvae_deconv3_gemm"),
    "scene_gather": ("r2_gather", "scene_gather"),
    "readout_feature_pool": ("r2_readout", "readout_feature_pool"),
    "scene_cnn_conv3_tile_resident": ("r2_conv5_l3", None),
}
UNITS = {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}


def read(rep):
    out = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(out.splitlines()))
    h, u, v = rows[0], rows[1], rows[2]
    g = lambda k: (float(v[h.index(k)].replace(",", "")), u[h.index(k)])
    tp = g("sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active")[0]
    rd, ru = g("dram__bytes_read.sum")
    wr, wu = g("dram__bytes_write.sum")
    dur, du = g("gpu__time_duration.sum")
    return {"tensor_pipe_pct": tp, "dram_bytes": rd * UNITS.get(ru, 1) + wr * UNITS.get(wu, 1),
            "duration_us": dur * {"ns": 1e-3, "us": 1, "ms": 1e3, "s": 1e6}.get(du, 1), "kernel": v[h.index("Kernel Name")][:80]}


def main():
    d = sys.argv[1] if len(sys.argv) > 1 else os.path.join(ROOT, "gpurun_out")
    tp_path, tr_path = os.path.join(ROOT, "profiles", "tensor_pipe.json"), os.path.join(ROOT, "profiles", "traffic.json")
    tp = json.load(open(tp_path)) if os.path.exists(tp_path) else {}
    tr = json.load(open(tr_path)) if os.path.exists(tr_path) else {}
    src = {}
    for entry, (tag, bench_name) in MAP.items():
        rep = os.path.join(d, "ncu_%s.ncu-rep" % tag)
        if not os.path.exists(rep):
            src[entry] = "stale (no %s)" % os.path.basename(rep)
            continue
        m = read(rep)
        if m["tensor_pipe_pct"] > 0:
            tp[entry] = round(m["tensor_pipe_pct"], 1)
        if bench_name:
            tr[bench_name] = int(m["dram_bytes"])
        src[entry] = "%s: %s, %.1f us under ncu" % (os.path.basename(rep), m["kernel"], m["duration_us"])
    tp["_note"] = ("sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active of single cold launches captured with "
                   "ncu --set full; written by tools/regen_profile_json.py, not by hand; not bench values")
    tp["_source"] = src
    tp["target_gru_pct"] = 70.0
    tr["_note"] = ("dram__bytes_read.sum + dram__bytes_write.sum per launch from the same captures (cold L2); keyed by "
                   "bench.py kernel name; written by tools/regen_profile_json.py")
    json.dump(tp, open(tp_path, "w"), indent=1)
    json.dump(tr, open(tr_path, "w"), indent=1)
    print(json.dumps(tp, indent=1))


if __name__ == "__main__":
    main()
