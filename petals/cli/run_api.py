from petals_b200.cli.run_api import main

if __name__ == "__main__":
    main()
