for v in eam morse "morse close" both oracle; do echo "== $v"; python tools/_smoke_dbg.py $v 2>&1 | tail -12; echo "rc=$?"; done
git -C . log --oneline | head -1
