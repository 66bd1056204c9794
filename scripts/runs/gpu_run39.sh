for m in 3 0 1; do for s in 2 4 8; do timeout 15 ./build_micro/mma_pipe $m $s || echo "mode $m stages $s: rc $?"; done; done
