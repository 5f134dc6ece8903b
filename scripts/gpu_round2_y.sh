# split launches (pair engine for the wide columns, full-size CTAs for the narrow tail) and the diagonal-block choice of the
# full-size CTA in batched launches
mkdir -p gpurun_out/r2y
for tw in 0 48 100 148 200 296; do
  echo "== IAK3D_TAIL_WIDTH=$tw"
  IAK3D_TAIL_WIDTH=$tw python scripts/stage_probe.py 5000 9 2>&1 | head -1
  IAK3D_TAIL_WIDTH=$tw python scripts/stage_probe.py 4000 36 2>&1 | head -1
  IAK3D_TAIL_WIDTH=$tw python scripts/stage_probe.py 2049 9 2>&1 | head -1
done 2>&1 | tee gpurun_out/r2y/tail_width.txt
for dg in 0 1; do
  echo "== IAK3D_DIAG=$dg"
  IAK3D_DIAG=$dg python scripts/stage_probe.py 1200 7 2>&1 | head -1
  IAK3D_DIAG=$dg python scripts/stage_probe.py 2049 9 2>&1 | head -1
  IAK3D_DIAG=$dg python scripts/stage_probe.py 3000 7 2>&1 | head -1
done 2>&1 | tee gpurun_out/r2y/diag.txt
