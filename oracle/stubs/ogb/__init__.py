"""TEST INFRASTRUCTURE ONLY. Stand-in for ogb==1.2.6 (requirement.yml:75)."""
