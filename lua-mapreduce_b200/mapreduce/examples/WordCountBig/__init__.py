"""The reference's WordCountBig example (mapreduce/examples/WordCountBig/): only the taskfn differs from
WordCount -- one map job per file of a directory of text splits."""
