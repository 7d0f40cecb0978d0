timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -5
for wl in ne30x72x40 ne256x128x10; do for m in 1 0; do echo $wl transposed=$m; CEDR_B200_TRANSPOSED=$m python tools/time_run.py qlt $wl; done; done
for nt in 8 16 32 64; do for m in 1 0; do echo nt=$nt transposed=$m; CEDR_B200_TRANSPOSED_MIN=1 CEDR_B200_TRANSPOSED=$m python tools/time_run.py qlt ne120x128x40 $nt; done; done
