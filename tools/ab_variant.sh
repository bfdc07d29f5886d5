# A/B of two builds of the library on the same box: make -C hept_b200/csrc VARIANT=<name> EXTRA=-D... first; usage: ab_variant.sh <name>
cd "$(dirname "$0")/.."
v=${1:-base}
for lib in "" _$v "" _$v; do
  echo "--- libhept_sm100$lib.so"
  HEPT_LIB=hept_b200/libhept_sm100$lib.so python tools/stage_times.py 60000 2>/dev/null | tail -n 1 | python -c "import json,sys; d=json.loads(sys.stdin.read()); print({k:d[k] for k in ('argsort','hat+tiles_fwd','fwd_call','bwd_call','bwd_pre+tiles')})"
done
