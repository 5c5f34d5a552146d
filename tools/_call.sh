python tools/_k2_lines.py
