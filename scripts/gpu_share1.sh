run() { echo -n "share=$2: "; RECOM_B200_SHARE=$2 timeout 100 python scripts/dev/quick.py $1 2>&1 | tail -1; echo; }
run "config2 --S 512 --reps 3" 1
run "config3 --S 512 --reps 3" 1
run "config3 --S 32 --reps 3" 1
run "config3 --S 512 --reps 3" 0
