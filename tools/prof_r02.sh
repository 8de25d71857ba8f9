#!/bin/bash
# Round-2 ncu evidence (run under gpurun, 1 GPU).  (1) every launch of two headline denoise steps with its device
# time and DRAM bytes; (2) ncu --set full of the dominant kernels at the headline shapes.  Outputs -> gpurun_out/.
set -u
mkdir -p gpurun_out
CMD="python bench.py --steps 2 --warmup 1 --no-tree --no-eager --no-cpu-baseline --no-text"
# warm-up step + e2e etc. come later in bench.py; the first ~450 launches after pipeline build are the warm-up step,
# the next ~900 the two timed steps: capture a window that covers them
# (only this library's kernels: the synthetic-weight initialisation launches thousands of torch kernels first)
KN='regex:^(gemm2_kernel|gemm_kernel|attn_kernel|ln_modulate_kernel|lora_down_kernel|lora_down_side_kernel|gemv_kernel|select_row_kernel|euler_step_kernel|advance_step_kernel|add2_kernel|add3_kernel|timestep_embed_kernel|f32_to_bf16_kernel)$'
ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none \
    -k "$KN" -c 1800 --csv --log-file gpurun_out/r02_launches.csv $CMD > gpurun_out/r02_launches.log 2>&1
echo "launch list rc=$?"
# full captures: attention (entry-B geometry), the mixed LoRA gate+res GEMM, the LoRA down-projection
ncu --set full --clock-control none --import-source on -k regex:attn_kernel -s 60 -c 2 \
    -o gpurun_out/r02_prof_attn $CMD > gpurun_out/r02_prof_attn.log 2>&1
echo "attn rc=$?"
ncu --set full --clock-control none --import-source on -k regex:gemm2_kernel -s 400 -c 8 \
    -o gpurun_out/r02_prof_gemm $CMD > gpurun_out/r02_prof_gemm.log 2>&1
echo "gemm rc=$?"
ncu --set full --clock-control none --import-source on -k regex:lora_down -s 60 -c 4 \
    -o gpurun_out/r02_prof_lora_down $CMD > gpurun_out/r02_prof_lora_down.log 2>&1
echo "lora_down rc=$?"
ncu --set full --clock-control none --import-source on -k regex:ln_modulate -s 60 -c 2 \
    -o gpurun_out/r02_prof_ln $CMD > gpurun_out/r02_prof_ln.log 2>&1
echo "ln rc=$?"
