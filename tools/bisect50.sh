#!/bin/bash
for rep in 1 2; do
  PYTHONPATH=_old_bb0f04a python tools/r50_graph.py 50 8 2>&1 | tail -1
  PYTHONPATH=. python tools/r50_graph.py 50 8 2>&1 | tail -1
  PYTHONPATH=_old_bb0f04a python tools/r50_graph.py 18 16 2>&1 | tail -1
  PYTHONPATH=. python tools/r50_graph.py 18 16 2>&1 | tail -1
done
