# A/B of attention_tc_kernel variants on one box: bash tools/ab_attention.sh tag1 tag2 ...   ("" = the in-tree library)
for r in 1 2; do
for v in "$@"; do
  if [ "$v" = "base" ]; then unset OT_B200_LIB; else export OT_B200_LIB=$PWD/onnx-transformer_b200/build/libot_b200_$v.so; fi
  echo "== variant [$v]"; python tools/bench_attention.py 2>&1 | tail -3 | head -2
done; done
