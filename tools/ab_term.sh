for i in 1 2; do
echo "== default"; python tools/time_terminal.py 1000000 120 3 | head -1
echo "== maxl1"; EMB_TERM_MAXL1=1 python tools/time_terminal.py 1000000 120 3 | head -1
done
