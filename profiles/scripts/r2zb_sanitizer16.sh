set -x
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
for tool in racecheck synccheck memcheck; do
  timeout 700 compute-sanitizer --tool $tool --print-limit 10 python tests/sanitizer_driver.py 16 3000 > gpurun_out/sanitizer4_${tool}_16cta.log 2>&1
  echo "rc=$?" >> gpurun_out/sanitizer4_${tool}_16cta.log
  grep -v "Host Frame\|Device Frame" gpurun_out/sanitizer4_${tool}_16cta.log | tail -5
done
