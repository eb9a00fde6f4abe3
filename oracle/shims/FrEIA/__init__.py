"""Import-only stub for FrEIA (absent).  models/SNF.py imports it at module
scope; only `energy_grad` (models/SNF.py:234-237) is ever used by the hot
path, and it needs nothing from FrEIA.  Test infrastructure only."""
