for np in 2 3; do
CEDR_B200_RING_NP=$np timeout 300 python tools/ring_check.py --skip-check --time ne120x128x40 --nt 1280 2>&1 | grep "ring=True"
done
python tools/ring_trace.py caas 86400 640 2>&1 | grep -E "run:|P |L |T |S |period"
