#!/bin/bash
# One GPU-box session: parity checks, bench lines, ncu launch list + full captures of the hot kernels.
# usage: bash tools/gpu_round.sh <tag> [steps...]   (steps: diag bench varlen qlora launches ncu_attn ncu_gemm pytest)
set -u
TAG=${1:-run}; shift || true
STEPS=${*:-"diag bench varlen qlora launches ncu_attn ncu_gemm"}
OUT=gpurun_out/$TAG
mkdir -p $OUT
nvidia-smi --query-gpu=name,clocks.max.sm,power.limit --format=csv > $OUT/gpu.txt 2>&1
for s in $STEPS; do
  case $s in
    diag) timeout 1500 python tools/gpu_diag.py ${DIAG_NAMES:-} > $OUT/diag.log 2>&1; cp gpurun_out/diag.json $OUT/diag.json 2>/dev/null;;
    timing) DTX_LIB_PATH=$PWD/datatunerx_b200/libdtxtune_timing.so timeout 300 python tools/attn_timing.py > $OUT/attn_timing.log 2>&1;;
    ncu_attn_bwd) timeout 900 ncu --set full --clock-control none --import-source on -k regex:attn_d --launch-skip 192 --launch-count 2 -o $OUT/attn_bwd -f \
                python bench.py --steps 1 --warmup 3 --no-cpu-baseline > $OUT/ncu_attn_bwd.log 2>&1;;
    sweep_gm) for g in 8 16 32; do DTX_GROUP_M=$g timeout 300 python bench.py --steps 6 --warmup 3 --no-cpu-baseline > $OUT/bench_gm$g.json 2> $OUT/bench_gm$g.err; done;;
    bench_fma) for e in 0 4 3; do DTX_FWD_EXP_FMA=$e timeout 300 python bench.py --steps 6 --warmup 3 --no-cpu-baseline > $OUT/bench_fma$e.json 2> $OUT/bench_fma$e.err; done;;
    attn_events) timeout 300 python tools/attn_timing.py > $OUT/attn_events.log 2>&1;;
    attn_launches) timeout 600 ncu --metrics gpu__time_duration.sum,sm__cycles_elapsed.max --clock-control none -k regex:attn_ --csv --log-file $OUT/attn_launches.csv python tools/attn_timing.py > $OUT/attn_launches.log 2>&1;;
    sanitizer) for c in gemm_nt gemm_kext rmsnorm cross_entropy adamw attn_fwd attn_bwd attn_varlen trainer_tiny; do
                 timeout 900 compute-sanitizer --tool racecheck --print-limit 5 python -m tests.gpu_checks $c > $OUT/racecheck_$c.log 2>&1; echo "$c rc=$?" >> $OUT/sanitizer_summary.txt; done
               timeout 900 compute-sanitizer --tool memcheck --print-limit 5 python -m tests.gpu_checks attn_varlen trainer_varlen > $OUT/memcheck_varlen.log 2>&1; echo "memcheck_varlen rc=$?" >> $OUT/sanitizer_summary.txt;;
    mgpu_check) timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node ${NGPU:-2} --master-addr 127.0.0.1 --master-port 29541 tests/multi_gpu_check.py > $OUT/multi_gpu_check_n${NGPU:-2}.log 2>&1;;
    mgpu_full) timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node ${NGPU:-2} --master-addr 127.0.0.1 --master-port 29544 tests/multi_gpu_check.py --full > $OUT/multi_gpu_check_full_n${NGPU:-2}.log 2>&1;;
    mgpu_ragged) timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node ${NGPU:-2} --master-addr 127.0.0.1 --master-port 29548 tests/multi_gpu_check.py --ragged > $OUT/multi_gpu_check_ragged_n${NGPU:-2}.log 2>&1;;
    mgpu_packed) timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node ${NGPU:-2} --master-addr 127.0.0.1 --master-port 29549 tests/multi_gpu_check.py --packed > $OUT/multi_gpu_check_packed_n${NGPU:-2}.log 2>&1;;
    bench_n) timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node ${NGPU:-2} --master-addr 127.0.0.1 --master-port 29542 bench.py --gpus ${NGPU:-2} --steps 8 --warmup 3 > $OUT/bench_n${NGPU:-2}.json 2> $OUT/bench_n${NGPU:-2}.err;;
    qlora_n) timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node ${NGPU:-2} --master-addr 127.0.0.1 --master-port 29543 bench.py --config mistral7b_qlora --gpus ${NGPU:-2} --steps 5 --warmup 3 > $OUT/bench_qlora_n${NGPU:-2}.json 2> $OUT/bench_qlora_n${NGPU:-2}.err;;
    full13b_n) timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node ${NGPU:-8} --master-addr 127.0.0.1 --master-port 29545 bench.py --config 13b_full --gpus ${NGPU:-8} --steps 6 --warmup 3 > $OUT/bench_13b_full_n${NGPU:-8}.json 2> $OUT/bench_13b_full_n${NGPU:-8}.err;;
    full13b_ctas) NCCL_MAX_CTAS=${NCCL_CTAS:-8} timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node ${NGPU:-8} --master-addr 127.0.0.1 --master-port 29547 bench.py --config 13b_full --gpus ${NGPU:-8} --steps 6 --warmup 3 > $OUT/bench_13b_full_n${NGPU:-8}_ctas${NCCL_CTAS:-8}.json 2> $OUT/bench_13b_full_n${NGPU:-8}_ctas${NCCL_CTAS:-8}.err;;
    small_full) timeout 600 python bench.py --config small_full --steps 6 --warmup 3 > $OUT/bench_small_full.json 2> $OUT/bench_small_full.err;;
    small_full_n) timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node ${NGPU:-2} --master-addr 127.0.0.1 --master-port 29546 bench.py --config small_full --gpus ${NGPU:-2} --steps 6 --warmup 3 > $OUT/bench_small_full_n${NGPU:-2}.json 2> $OUT/bench_small_full_n${NGPU:-2}.err;;
    jobs_tiny) timeout 600 python tools/concurrent_jobs.py --jobs ${NJOBS:-1} --gpus-per-job 2 --model tiny --steps 12 --out $OUT/concurrent_jobs_tiny.json > $OUT/concurrent_jobs_tiny.log 2>&1;;
    jobs_7b) timeout 900 python tools/concurrent_jobs.py --jobs ${NJOBS:-4} --gpus-per-job 2 --model 7b --steps ${JOB_STEPS:-24} --out $OUT/concurrent_jobs_7b.json > $OUT/concurrent_jobs_7b.log 2>&1;;
    jobs_ragged) for v in 0 1; do DTX_OPTIONS=varlen_split=$v timeout 900 python tools/concurrent_jobs.py --jobs 1 --gpus-per-job ${NGPU:-1} --model 7b --steps ${JOB_STEPS:-12} --ragged --out $OUT/worker_ragged_split$v.json > $OUT/worker_ragged_split$v.log 2>&1; done;;
    memcheck_packed) timeout 900 compute-sanitizer --tool memcheck --print-limit 5 python -m tests.gpu_checks trainer_varlen > $OUT/memcheck_varlen_packed.log 2>&1
                     timeout 900 compute-sanitizer --tool racecheck --print-limit 5 python -m tests.gpu_checks trainer_varlen > $OUT/racecheck_varlen_packed.log 2>&1;;
    memcheck_groups) timeout 900 compute-sanitizer --tool memcheck --print-limit 5 python -m tests.gpu_checks trainer_varlen_groups > $OUT/memcheck_varlen_groups.log 2>&1
                     timeout 900 compute-sanitizer --tool racecheck --print-limit 5 python -m tests.gpu_checks trainer_varlen_groups > $OUT/racecheck_varlen_groups.log 2>&1;;
    determinism) timeout 600 python tools/diag_determinism.py > $OUT/determinism.log 2>&1;;
    step_repeat) timeout 600 python tools/diag_step_repeat.py > $OUT/step_repeat.log 2>&1;;
    initcheck) timeout 900 compute-sanitizer --tool initcheck --print-limit 20 python -m tests.gpu_checks trainer_tiny > $OUT/initcheck_trainer_tiny.log 2>&1;;
    load_paths) timeout 600 python tools/diag_load_paths.py > $OUT/load_paths.log 2>&1;;
    diag7b) timeout 600 python tools/diag_layer7b.py > $OUT/diag7b.log 2>&1;;
    parity7b) timeout 1200 python tools/parity_7b.py --steps 3 --out $OUT/parity_7b.json > $OUT/parity_7b.log 2>&1;;
    pytest) timeout 1500 python -m pytest tests -m gpu -x -q > $OUT/pytest_gpu.log 2>&1;;
    smoke) timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > $OUT/smoke.log 2>&1;;
    bench) timeout 600 python bench.py --steps 8 --warmup 3 --no-cpu-baseline > $OUT/bench.json 2> $OUT/bench.err;;
    benchcpu) timeout 900 python bench.py --steps 8 --warmup 3 > $OUT/bench_cpu.json 2> $OUT/bench_cpu.err;;
    refarm) timeout 900 python bench.py --impl reference --steps 20 --warmup 5 > $OUT/bench_reference.json 2> $OUT/bench_reference.err;;
    varlen) timeout 600 python bench.py --config 7b_varlen --steps 8 --warmup 3 > $OUT/bench_varlen.json 2> $OUT/bench_varlen.err;;
    varlen_cost) for v in ${COSTS:-250 500 1000}; do DTX_VARLEN_GROUP_COST=$v timeout 600 python bench.py --config 7b_varlen --steps 8 --warmup 3 > $OUT/bench_varlen_cost$v.json 2> $OUT/bench_varlen_cost$v.err; done;;
    varlen_n) timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node ${NGPU:-8} --master-addr 127.0.0.1 --master-port 29550 bench.py --config 7b_varlen --gpus ${NGPU:-8} --steps 8 --warmup 3 > $OUT/bench_varlen_n${NGPU:-8}.json 2> $OUT/bench_varlen_n${NGPU:-8}.err;;
    varlen_pack) for v in 0 1; do DTX_VARLEN_PACK=$v timeout 600 python bench.py --config 7b_varlen --steps 8 --warmup 3 > $OUT/bench_varlen_pack$v.json 2> $OUT/bench_varlen_pack$v.err; done;;
    varlen_ab) for v in 0 1; do DTX_VARLEN_SPLIT=$v timeout 600 python bench.py --config 7b_varlen --steps 8 --warmup 3 > $OUT/bench_varlen_split$v.json 2> $OUT/bench_varlen_split$v.err; done;;
    qlora) timeout 600 python bench.py --config mistral7b_qlora --steps 5 --warmup 3 > $OUT/bench_qlora.json 2> $OUT/bench_qlora.err;;
    launches) timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none --launch-skip 1600 -c 2000 --csv --log-file $OUT/launches.csv \
                python bench.py --steps 1 --warmup 3 --no-cpu-baseline > $OUT/launches_bench.log 2>&1;;
    ncu_gemm_bwd) timeout 900 ncu --set full --clock-control none --import-source on -k regex:gemm2_kernel --launch-skip 660 --launch-count 10 -o $OUT/gemm_bwd -f \
                python bench.py --steps 1 --warmup 3 --no-cpu-baseline > $OUT/ncu_gemm_bwd.log 2>&1;;
    ncu_attn3) timeout 900 ncu --set full --clock-control none --import-source on -k regex:attn_ --launch-skip 318 --launch-count 4 -o $OUT/attn3 -f \
                python bench.py --steps 1 --warmup 3 --no-cpu-baseline > $OUT/ncu_attn3.log 2>&1;;
    qlora_ab) for pf in 1 0; do DTX_NF4_PREFETCH=$pf timeout 600 python bench.py --config mistral7b_qlora --steps 6 --warmup 3 > $OUT/bench_qlora_prefetch$pf.json 2> $OUT/bench_qlora_prefetch$pf.err; done;;
    ncu_attn) timeout 900 ncu --set full --clock-control none --import-source on -k regex:attn_ --launch-skip 12 --launch-count 3 -o $OUT/attn -f \
                python bench.py --steps 1 --warmup 3 --no-cpu-baseline > $OUT/ncu_attn.log 2>&1;;
    ncu_gemm) timeout 900 ncu --set full --clock-control none --import-source on -k regex:gemm2_kernel --launch-skip 40 --launch-count 12 -o $OUT/gemm -f \
                python bench.py --steps 1 --warmup 3 --no-cpu-baseline > $OUT/ncu_gemm.log 2>&1;;
    *) echo "unknown step $s";;
  esac
  echo "$s done rc=$? $(date +%T)" >> $OUT/steps.log
done
tail -3 $OUT/diag.log 2>/dev/null
cat $OUT/bench.json 2>/dev/null | head -c 1500
