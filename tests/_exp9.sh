cd /root/repo
for tool in memcheck racecheck synccheck; do
  echo "== compute-sanitizer --tool $tool"
  timeout 900 compute-sanitizer --tool $tool python tests/diag_sanitize.py 2>&1 | grep -v "^$" | tail -12
done
