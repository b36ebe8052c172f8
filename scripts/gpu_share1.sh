run() { echo -n "$1 share=$3: "; RECOM_B200_SHARE=$3 RECOM_B200_LIB=$PWD/variants/lib_$1.so timeout 60 python scripts/dev/quick.py $2 2>&1 | tail -1; echo; }
for m in 0 1; do
run share3 "config3 --chains 1024 --S 32 --reps 2" $m
run share3 "config2 --chains 1024 --S 32 --reps 2" $m
run share3 "config3 --chains 64 --S 50 --reps 2" $m
run share3 "config3 --chains 1 --S 100 --reps 2" $m
run share3 "config3 --wilson --chains 1024 --S 32 --reps 2" $m
run share3 "config3 --S 512 --reps 2" $m
done
