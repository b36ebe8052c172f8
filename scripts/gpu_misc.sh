for SH in 0 3; do
  echo "== phases share=$SH config2"; RECOM_B200_SHARE=$SH RECOM_B200_HELP_AFTER=100000000 RECOM_B200_LIB=variants/lib_timers.so timeout 120 python scripts/phases.py --workload config2 --chain-steps 128 2>&1 | grep -v "wilson\|slot14" | tail -12
done
