# B200SR_EPI_DEBUG experiments on the slow (Cin or Cout = 64) layers: bits 1..64 switch parts of the epilogue off
# (results are then wrong: timing only); 128 = evict_last hint on activation loads, 256 = evict_first hint on output stores
for m in ${@:-0 127 128 256 384}; do echo "EPI_DEBUG=$m"; B200SR_EPI_DEBUG=$m python tools/bench_layers.py fwd dgrad 2>&1 | grep -E "enc1.3|enc2.0|enc2.3|dec1.0|total"; done
