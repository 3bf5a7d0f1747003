for ns in 0 4000 8000 16000 32000; do echo "stagger $ns"; POP_F2_STAGGER_NS=$ns TUNE_CS=2 TUNE_OCC=0 python tools/tune_f2.py 0,1 2>&1 | grep "fused-reg"; done
for ns in 0 4000 16000; do echo "stagger $ns"; POP_F2_STAGGER_NS=$ns TUNE_CS=1 TUNE_OCC=0 python tools/tune_f2.py 2,3 2>&1 | grep "fused-reg"; done
