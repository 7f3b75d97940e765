# usage: exp_lv.sh NAME...  -- per-kernel durations of one CCpol gradient pass for the in-tree library and each variant
bash tools/dev/lv.sh base 2>&1 | tail -9
for n in "$@"; do bash tools/dev/lv.sh $n $PWD/tools/dev/variants/libpimdk_$n.so 2>&1 | grep -E "setup|dipind|total"; done
