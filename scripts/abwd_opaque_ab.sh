# same-box A/B of two builds of the library (PM_B200_LIB): attention backward with / without the opaque per-thread constants
# (the plain build: PM_NVCC_EXTRA=-DPM_ABWD_PLAIN python -m paintmind_b200.build, copied to lib/libpaintmind_b200_plain.so)
PLAIN=$PWD/paintmind_b200/lib/libpaintmind_b200_plain.so
for v in plain opaque plain opaque; do
  if [ $v = plain ]; then export PM_B200_LIB=$PLAIN; else unset PM_B200_LIB; fi
  echo "$v: $(python scripts/bringup_bwd.py full attn 2>&1 | grep 'attn bwd B=256\|FAIL' | tr '\n' ' ')"
done
unset PM_B200_LIB
