# State digests of C4 under different stream configurations (all must be equal): sh tools/det_modes.sh
for m in 31 15 27 23 29 30 0; do SYMGPU_SIDE=$m python tools/gpu_determinism.py 60 5 > gpurun_out/det_m$m.log 2>&1; done
CUDA_LAUNCH_BLOCKING=1 python tools/gpu_determinism.py 60 5 > gpurun_out/det_blk.log 2>&1
for f in m31 m15 m27 m23 m29 m30 m0 blk; do echo $f $(awk '{print $NF}' gpurun_out/det_$f.log | tr '\n' ' ' | cut -c1-215); done
