# same-box A/B of two builds of the library for the forward attention kernel (PM_B200_LIB=<other build>): isolated burst + parity
OTHER=$PWD/paintmind_b200/lib/${1:-libpaintmind_b200_early.so}
for v in base other base other; do
  if [ $v = other ]; then export PM_B200_LIB=$OTHER; else unset PM_B200_LIB; fi
  echo "$v: $(python scripts/attn3_ab.py w16:1 2>&1 | tail -1)"
done
unset PM_B200_LIB
