"""Stand-in for netcal (lib/metrics/utils.py:16); only the name ECE is needed."""
