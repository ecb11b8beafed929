"""CPU oracle for the FPS hot path.  TEST INFRASTRUCTURE ONLY -- never imported by fpsample_b200."""
