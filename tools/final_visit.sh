bash tools/gpu_check.sh all r02i
out=gpurun_out/r02i
B200BO_PREDICT_IMPL=tf32 timeout 600 ncu --set full --clock-control none -k regex:predict_acq_tc4 -s 3 -c 1 -o $out/prof_predict_tc4 -f python bench.py --steps 1 --warmup 3 --no-cpu-baseline --no-extra > $out/ncu_full_tc4.log 2>&1
ncu -i $out/prof_predict_tc4.ncu-rep --page raw --csv 2>/dev/null | gzip > $out/prof_predict_tc4.raw.csv.gz; rm -f $out/prof_predict_tc4.ncu-rep
B200BO_GRAPH=0 timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file $out/launches_lml.csv python tools/lml_once.py > $out/lml_once.log 2>&1; gzip -f $out/launches_lml.csv
REPS=1 timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file $out/launches_suggest.csv python tools/suggest_once.py > $out/suggest_once.log 2>&1; gzip -f $out/launches_suggest.csv
ls -la $out | tail -12
