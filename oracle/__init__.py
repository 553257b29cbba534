"""CPU oracle for the hybrid-retrieval hot path -- TEST INFRASTRUCTURE ONLY (see exact_scan.c header)."""
