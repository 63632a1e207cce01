D=gpurun_out/san
mkdir -p $D
for tool in memcheck racecheck initcheck; do
  timeout 400 compute-sanitizer --tool $tool python profiles/sanitize_driver.py > $D/$tool.log 2>&1
  grep -E "SUMMARY|done," $D/$tool.log | tail -2
done
