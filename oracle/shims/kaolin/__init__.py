"""Shim: stands in for NVIDIAGameWorks/kaolin (absent here).  Test infrastructure only."""
