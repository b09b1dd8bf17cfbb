for rep in 1 2 3; do
echo "--- new"; python tools/kbench.py --games 1048576 4194304 --reps 5 --rollout 64 | tail -2
echo "--- old"; ORX_LIB=$PWD/optimax_rogue_b200/liborx_old.so python tools/kbench.py --games 1048576 4194304 --reps 5 --rollout 64 | tail -2
done
