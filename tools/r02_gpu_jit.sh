#!/bin/bash
# round 2: schedule fuzzing -- the parity tests of the hand-over protocol and the pipeline under the jitter build
# (-DORX_PIPE_JITTER: pseudo-random pauses of up to 2 us wherever producer, consumers and consecutive launches hand over)
mkdir -p gpurun_out
{
echo "# ORX_LIB=liborx_jitter.so (tools/build_variant.py jitter -DORX_PIPE_JITTER): $(cuobjdump -sass optimax_rogue_b200/liborx_jitter.so | grep -c NANOSLEEP) NANOSLEEP sites; the shipped library has $(cuobjdump -sass optimax_rogue_b200/liborx.so | grep -c NANOSLEEP)"
for round in 1 2; do
  echo "=== round $round"
  ORX_LIB=$PWD/optimax_rogue_b200/liborx_jitter.so timeout 1500 python -m pytest tests/test_gpu_tile_flags.py tests/test_gpu_parity.py tests/test_gpu_r1.py tests/test_gpu_guard_bytes.py -m gpu -q -p no:cacheprovider 2>&1 | tail -n 6
done
echo "=== fuzz campaign under jitter (90 s)"
ORX_LIB=$PWD/optimax_rogue_b200/liborx_jitter.so timeout 200 python tests/fuzz_campaign.py 90 2>&1 | head -n 2
} > gpurun_out/jitter_tests.log 2>&1; cat gpurun_out/jitter_tests.log
