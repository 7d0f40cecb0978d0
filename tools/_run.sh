set -x
timeout 300 python tools/ring_check.py --time ne120x128x40 --nt 1280 > gpurun_out/ring2.log 2>&1; echo exit $?
for np in 2 3; do for sw in 1 2; do
CEDR_B200_RING_NP=$np CEDR_B200_RING_SW=$sw timeout 300 python tools/ring_check.py --skip-check --time ne120x128x40 --nt 1280 2>&1 | grep "ring=True" >> gpurun_out/ring2.log
done; done
timeout 300 python tools/ring_check.py --skip-check --time ne30x72x40 >> gpurun_out/ring2.log 2>&1
grep -E "TIME|OK|FAIL|MISM|gave" gpurun_out/ring2.log
