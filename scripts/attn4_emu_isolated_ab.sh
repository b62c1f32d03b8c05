for v in 1 2 0 1 2 0; do
  echo "emu $v isolated: $(PM_ATTN4_VARIANT=$v python scripts/attn3_ab.py w16:$v 2>&1 | tail -1)"
done
