cd /root/repo
for t in test_summary_ext test_odd test_moments_mode test_checkpoint test_global_ids test_gathered test_synth_reports test_two_frame; do
  echo "=== $t"; timeout 300 python -m pytest tests/test_gpu_round2.py -k $t -q -x 2>&1 | grep -v "^  File\|^Extension" | tail -25
done
