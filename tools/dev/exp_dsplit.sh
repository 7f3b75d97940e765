V=$PWD/tools/dev/variants/libpimdk_dsplit.so
timeout 200 python -m pytest tests -m gpu -x -q 2>&1 | tail -1
PIMDK_LIB=$V timeout 200 python -m pytest tests -m gpu -x -q 2>&1 | tail -1
bash tools/dev/lv.sh base 2>&1 | tail -9
bash tools/dev/lv.sh dsplit $V 2>&1 | tail -9
b() { timeout 100 python bench.py --quick --steps 4 --warmup 3 2>/dev/null | python -c "import sys,json; d=json.loads(sys.stdin.read()); print('$1', d['ms_per_step'], d['value'])"; }
b base; PIMDK_LIB=$V b dsplit; b base; PIMDK_LIB=$V b dsplit
