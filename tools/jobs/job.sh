for v in 15 0; do
echo "=== EPI stamps (col1 = loop top, col2 = before tfull wait) ARL_FWD_WIDE=$v"
ARL_FWD_WIDE=$v ARL_LIB_PATH=/root/repo/accel_rl_b200/csrc/libaccelrl_b200_trace.so python tests/trace_persist.py 2>&1 | tail -17 | head -12
done
