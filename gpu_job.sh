python scripts/probe_f32.py
