# A/B of two builds of the library on the same box: make -C hept_b200/csrc VARIANT=base first (from the commit to compare with)
cd "$(dirname "$0")/.."
for v in "" _base "" _base; do
  echo "--- libhept_sm100$v.so"
  HEPT_LIB=hept_b200/libhept_sm100$v.so python tools/stage_times.py 60000 2>/dev/null | tail -n 1 | python -c "import json,sys; d=json.loads(sys.stdin.read()); print({k:d[k] for k in ('argsort','hat+tiles_fwd','fwd_call','bwd_call','prepare_input_batched')})"
done
