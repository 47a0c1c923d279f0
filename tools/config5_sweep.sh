# usage: config5_sweep.sh N   -> profiles-ready JSON lines in gpurun_out/r02_config5_n$N.jsonl
N=$1
OUT=gpurun_out/r02_config5_n$N.jsonl
: > $OUT
for mode in exact_tc bf16; do
  if [ "$N" = "1" ]; then
    python bench.py --config 5 --images 5000 --mode $mode >> $OUT 2>> gpurun_out/r02_config5_n$N.err
  else
    python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $N --config 5 --images 5000 --mode $mode --single-file >> $OUT 2>> gpurun_out/r02_config5_n$N.err
  fi
done
cat $OUT | cut -c1-600
