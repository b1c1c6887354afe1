"""Test/benchmark harness helpers (NOT product code): DreamScene-convention cameras and the
deterministic synthetic scenes of SURVEY.md section 8(d).  Used by tests/, bench.py, benchmarks/ and tools/."""
