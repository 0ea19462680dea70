"""mapreduce/examples/WordCount mirrored as Python plugin modules (see init.py)."""
