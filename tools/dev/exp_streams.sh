# usage: exp_streams.sh NAME...  -- bit-exactness of each variant library's multi-pass path, then C4 bench (quick) base vs variants
b() { timeout 100 python bench.py --quick --steps 4 --warmup 3 2>/dev/null | python -c "import sys,json; d=json.loads(sys.stdin.read()); print('$1', d['ms_per_step'], d['value'])"; }
for n in "$@"; do PIMDK_LIB=$PWD/tools/dev/variants/libpimdk_$n.so timeout 120 python tools/dev/check_streams.py 2>&1 | tail -1; done
for i in 1 2; do
b base
for n in "$@"; do PIMDK_LIB=$PWD/tools/dev/variants/libpimdk_$n.so b $n; done
done
