for kb in 0 44 56 75; do
  echo "== TRIG_SMEM_KB=$kb"
  SBD_TRIG_SMEM_KB=$kb PROBE_FLAGS=1 timeout 120 python scripts/probe_step.py c2 2>&1 | grep -E "wall|per-kernel" | cut -c1-120
done
