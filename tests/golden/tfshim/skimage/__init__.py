"""stub (tests/golden only)"""
