python -c "import __graft_entry__ as g; g.smoke()"; echo "smoke rc=$?"
python -m pytest tests/test_search.py tests/test_process_search.py -m gpu -x -q 2>&1 | tail -2
