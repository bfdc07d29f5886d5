# A/B of HEPT_BWD_SPLIT (attn_bwd_tc.cu): build the variant first: make -C hept_b200/csrc VARIANT=split EXTRA=-DHEPT_BWD_SPLIT=1
cd /root/repo
echo "--- default (no split)"; python tools/stage_times.py 60000 2>/dev/null | tail -n 1 | python -c "import json,sys; d=json.loads(sys.stdin.read()); print({k:d[k] for k in ('bwd_call','bwd_pre+tiles','fwd_call')})"
echo "--- split"; HEPT_LIB=hept_b200/libhept_sm100_split.so python tools/stage_times.py 60000 2>/dev/null | tail -n 1 | python -c "import json,sys; d=json.loads(sys.stdin.read()); print({k:d[k] for k in ('bwd_call','bwd_pre+tiles','fwd_call')})"
python -m pytest tests/test_gpu_parity.py -m gpu -q -x -k "backward or module or head_groups or clamp or other_block or engines" 2>&1 | tail -n 3
