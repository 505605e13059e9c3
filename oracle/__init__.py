"""Test infrastructure: CPU oracle of the kmerseek hot path. Never imported by kmerseek_b200/."""
