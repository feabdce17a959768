#!/bin/bash
# compute-sanitizer over a small, representative slice of the GPU suites (memcheck + racecheck).
# Usage (GPU box): bash tools/sanitize.sh  -> gpurun_out/sanitizer_*.log
mkdir -p gpurun_out
SEL='test_single_ranges_all_alignments_R8 or test_coalesced_identical_gets or test_wave_of_mixed_requests or test_verify_flag_and_corruption_X4 or test_crc32c_device_kernel_K3'
timeout 1200 compute-sanitizer --tool memcheck --leak-check no --error-exitcode 9 python -m pytest tests/test_gpu_parity.py -m gpu -q -x -k "$SEL" > gpurun_out/sanitizer_memcheck.log 2>&1
echo "memcheck rc=$?" >> gpurun_out/sanitizer_memcheck.log
timeout 1200 compute-sanitizer --tool memcheck --leak-check no --error-exitcode 9 python -m pytest tests/test_gpu_inflate.py -m gpu -q -x -k "test_gzip_header_fields or test_declined_streams_G4 or test_inflate_to_host_and_ranges_G3" > gpurun_out/sanitizer_memcheck_inflate.log 2>&1
echo "memcheck rc=$?" >> gpurun_out/sanitizer_memcheck_inflate.log
timeout 1200 compute-sanitizer --tool racecheck --error-exitcode 9 python -m pytest tests/test_gpu_parity.py -m gpu -q -x -k "test_coalesced_identical_gets or test_crc32c_device_kernel_K3" > gpurun_out/sanitizer_racecheck.log 2>&1
echo "racecheck rc=$?" >> gpurun_out/sanitizer_racecheck.log
for f in gpurun_out/sanitizer_*.log; do tail -n 4 "$f"; done
# fused verify (K2+K3): shared-memory tile checksums, mbarrier waits by every thread
SELV='test_whole_objects_verify_clean_every_size or test_destination_phase_and_ranges or test_coalesced_group_mixed_flags_and_corruption'
timeout 1500 compute-sanitizer --tool memcheck --leak-check no --error-exitcode 9 python -m pytest tests/test_gpu_verify_fused.py -m gpu -q -x -k "$SELV" > gpurun_out/sanitizer_memcheck_verify.log 2>&1
echo "memcheck rc=$?" >> gpurun_out/sanitizer_memcheck_verify.log
timeout 1500 compute-sanitizer --tool racecheck --error-exitcode 9 python -m pytest tests/test_gpu_verify_fused.py -m gpu -q -x -k "test_whole_objects_verify_clean_every_size or test_destination_phase_and_ranges" > gpurun_out/sanitizer_racecheck_verify.log 2>&1
echo "racecheck rc=$?" >> gpurun_out/sanitizer_racecheck_verify.log
for f in gpurun_out/sanitizer_*verify.log; do tail -n 4 "$f"; done
