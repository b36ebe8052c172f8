# A/B of library builds on ONE box with the quick driver: bash scripts/gpu_ab2.sh "<quick args>" tagA tagB ...
ARGS="$1"; shift
for rep in 1 2; do
  for tag in "$@"; do
    echo -n "$tag: "; RECOM_B200_LIB=$PWD/variants/lib_$tag.so python scripts/dev/quick.py $ARGS 2>&1 | tail -1
  done
done
