#!/bin/bash
# The UNMODIFIED reference trainer (DS_NeRF/run_nerf.py --no_tcnn) over the drop-in on a real GPU, by the recipe of INTEGRATION.md
# section 1.  The GPU box has no /root/reference: before the gpurun call, in the build container,
#     mkdir -p baseline/_ref && cp -r /root/reference/DS_NeRF baseline/_ref/DS_NeRF      (git-ignored, travels with gpurun)
# Usage: bash tools/run_reference_trainer_gpu.sh [iters] [tag]     -> gpurun_out/<tag>_reference_trainer.log (+ ncu launch list)
set -u
ITERS="${1:-60}"; TAG="${2:-rXX}"
REPO="$(cd "$(dirname "$0")/.." && pwd)"
REF="$REPO/baseline/_ref"
[ -f "$REF/DS_NeRF/run_nerf.py" ] || REF=/root/reference
[ -f "$REF/DS_NeRF/run_nerf.py" ] || { echo "no reference checkout (baseline/_ref/DS_NeRF or /root/reference)"; exit 2; }
OUT="$REPO/gpurun_out"; mkdir -p "$OUT"
WORK="$(mktemp -d /tmp/spn_ref_XXXX)"
python - "$WORK/scene" <<PY
import importlib, sys
sys.path.insert(0, "$REPO")
sio = importlib.import_module("spin-nerf_b200.scene_io")
info = sio.synthetic_scene(sys.argv[1], n_views=8, hw=(96, 128), factor=2, seed=0, n_unlabelled=1)
print("scene written:", sys.argv[1], {k: (getattr(v, "shape", v)) for k, v in info.items() if k in ("focal",)})
PY
cd "$WORK" && mkdir -p lama/LaMa_test_images
ARGS=(--expname t --datadir "$WORK/scene" --basedir "$WORK/logs" --dataset_type llff --factor 2 --N_rand 1024 --N_samples 64
      --N_importance 64 --use_viewdirs --raw_noise_std 1.0 --no_ndc --lindisp --white_bkgd --no_tcnn --N_gt 0 --i_video 100000
      --i_testset 100000 --i_weights 100000 --i_feat 100000 --i_print 10)
export PYTHONPATH="$REPO/spin-nerf_b200/dropin:$REPO/spin-nerf_b200/compat:$REF/DS_NeRF"
echo "=== unmodified $(md5sum "$REF/DS_NeRF/run_nerf.py" | cut -c1-12) run_nerf.py, $ITERS iterations, N_rand 1024, 64+64 samples"
( time timeout 600 python -P "$REF/DS_NeRF/run_nerf.py" "${ARGS[@]}" --N_iters "$ITERS" ) > "$OUT/${TAG}_reference_trainer.log" 2>&1
echo "rc=$?"; grep -E "TRAIN|Iter|real|Error|error" "$OUT/${TAG}_reference_trainer.log" | tail -12
# kernel list of ~8 iterations of the same command (cold-cache, serialised: shares only)
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -s 2000 -c 600 --csv --log-file "$OUT/${TAG}_reference_trainer_launches.csv" \
    python -P "$REF/DS_NeRF/run_nerf.py" "${ARGS[@]}" --N_iters 24 > "$OUT/${TAG}_reference_trainer_ncu.log" 2>&1
echo "ncu rc=$?"
