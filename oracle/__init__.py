"""CPU oracle for the PointsToWood hot path -- test infrastructure, never the product path."""
