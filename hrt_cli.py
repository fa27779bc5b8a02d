#!/usr/bin/env python
"""`python hrt_cli.py --scene final --width 800 --height 800 --samples 1000 --depth 50 --out final.png`
Entry point for the headless front end in hyper-ray-tracer_b200/__main__.py (same flags as the reference's CLI)."""
import importlib
import sys

import __graft_entry__ as graft

graft.load_package()
sys.exit(importlib.import_module(graft.PKG_NAME + ".__main__").main())
