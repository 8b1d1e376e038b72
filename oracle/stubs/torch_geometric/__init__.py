"""TEST INFRASTRUCTURE ONLY. Minimal pure-torch stand-in for torch-geometric==1.6.3
(requirement.yml:97) - see oracle/README.md. Never imported by the product package."""
