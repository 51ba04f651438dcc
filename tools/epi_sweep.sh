for m in 0 64 32 2 8 1 96 98 127; do echo "EPI_DEBUG=$m"; B200SR_EPI_DEBUG=$m python tools/bench_layers.py fwd dgrad 2>&1 | grep -E "enc1.3|enc2.0|dec1.0|total"; done
