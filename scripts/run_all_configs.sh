#!/usr/bin/env bash
# Reproduce profiles/<tag>_all_workloads: one bench.py line per BASELINE.json config on 1 GPU, the reference arm,
# and the training step (configs[3]).  Usage: scripts/run_all_configs.sh [out_dir] [tag]
# GD_SWEEP_FAST=1 skips the 15 s CPU baseline of the secondary workloads (the headline run keeps it).
set -u
EXTRA=""; [ -n "${GD_SWEEP_FAST:-}" ] && EXTRA="--no-cpu-baseline"
OUT=${1:-gpurun_out}; TAG=${2:-run}
mkdir -p "$OUT"
python bench.py > "$OUT/${TAG}_bench.json" 2> "$OUT/${TAG}_bench.err"
python bench.py --impl reference --steps 3 --warmup 1 > "$OUT/${TAG}_bench_reference.json" 2>> "$OUT/${TAG}_bench.err"
for w in cgnni_ldpc_awgn_B1024 cgnni_bch_awgn_B1024 cgnni_bch_awgn_B65536 \
         v2_4_toric_L5_iidxz_B65536 v2_4_toric_L11_iidxz_B65536 v2_4_rotated_d11_depol_B65536 \
         qgnni_hgp1600_depol_B16384_T50 bp_hgp1600_depol_B16384_T50 v2_4_hgp1600_depol_B8192; do
  python bench.py --workload "$w" --steps 10 $EXTRA > "$OUT/${TAG}_bench_$w.json" 2>> "$OUT/${TAG}_bench.err"
done
python scripts/train_bench.py 7 4096 > "$OUT/${TAG}_train_bench.txt" 2>&1
python - "$OUT" "$TAG" <<'PY'
import glob, json, sys
out, tag = sys.argv[1], sys.argv[2]
for f in sorted(glob.glob("%s/%s_bench*.json" % (out, tag))):
    try:
        d = json.loads(open(f).read().strip().splitlines()[-1])
        print("%-34s %-9s value %.4g  e2e %.4g  ms/step %.3f  hbm frac %s" % (
            d["config"]["workload"], d.get("impl", "ours"), d["value"], d["e2e"]["value"], d["ms_per_step"],
            d.get("roofline", {}).get("frac")))
    except Exception as e:
        print(f, "ERR", e)
PY
tail -2 "$OUT/${TAG}_train_bench.txt"
