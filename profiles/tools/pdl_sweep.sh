# A/B of programmatic dependent launch per kernel of the reverse walk: TACO_PDL is a bit mask (1 cell, 2 skinny_tc, 4 attention, 8 query)
for m in ${MASKS:-0 15 13}; do TACO_PDL=$m python profiles/tools/c5_kernel_times.py > gpurun_out/pdl_$m.log 2>&1; echo "mask $m: $(grep -m1 '^C5 step' gpurun_out/pdl_$m.log | cut -c1-75)"; done
