tools/ab.sh "main gs01 gs001 gs1e4" "2"
for v in gs01 gs001 gs1e4; do echo "== parity $v"; QPPVM_B200_LIB=$PWD/qppvm_b200/variants/libqppvm_b200_$v.so python -m pytest tests/test_gpu_parity.py tests/test_golden.py -m gpu -x -q 2>&1 | tail -3; done
