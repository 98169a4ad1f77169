"""TEST INFRASTRUCTURE (CPU oracle). Never imported by the product package."""
