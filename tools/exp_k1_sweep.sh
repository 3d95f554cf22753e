set -x
cd $GRAFT_REPO_ROOT
nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tools/k1_micro tools/k1_micro.cu
UBPL_AB_B=64,128,256,512 UBPL_AB_MASKS=0,16,4,11 UBPL_AB_CAPS=8 timeout 600 python tools/k1_ab.py c2 > gpurun_out/s2_k1_sweep_c2.log 2>&1
UBPL_AB_MASKS=0,16 UBPL_AB_CAPS=8 timeout 600 python tools/k1_ab.py c4 c3 > gpurun_out/s2_k1_tl_c4.log 2>&1
for n in 7168 14336 28672 57344; do echo "N=$n"; timeout 300 tools/k1_micro $n; done > gpurun_out/s2_micro_sweep.log 2>&1
tail -40 gpurun_out/s2_k1_sweep_c2.log
