"""see backbone/__init__.py"""
