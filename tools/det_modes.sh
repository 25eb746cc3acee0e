for m in 15 11 7 13 14 0; do SYMGPU_SIDE=$m python tools/gpu_determinism.py 60 5 > gpurun_out/det_m$m.log 2>&1; done
CUDA_LAUNCH_BLOCKING=1 python tools/gpu_determinism.py 60 5 > gpurun_out/det_blk.log 2>&1
CUDA_LAUNCH_BLOCKING=1 SYMGPU_SIDE=0 python tools/gpu_determinism.py 60 5 > gpurun_out/det_blk0.log 2>&1
for f in m15 m11 m7 m13 m14 m0 blk blk0; do echo $f $(awk '{print $NF}' gpurun_out/det_$f.log | tr '\n' ' ' | cut -c1-215); done
