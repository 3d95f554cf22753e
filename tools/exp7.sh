cd $GRAFT_REPO_ROOT
UBPL_AB_EMA=1 UBPL_AB_MASKS=0,16 UBPL_AB_CAPS=8 timeout 600 python tools/k1_ab.py c2 > gpurun_out/s2_k1_ema_tl.log 2>&1
cat gpurun_out/s2_k1_ema_tl.log
UBPL_AB_EMA=0 UBPL_AB_MASKS=0,16 UBPL_AB_CAPS=8 timeout 600 python tools/k1_ab.py c2
