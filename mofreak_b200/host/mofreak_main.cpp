// Command-line driver in the shape of the reference's main.cpp stages (main.cpp:854-924 computeMoFREAKFiles,
// :294-349 computeBOWRepresentation), with paths as arguments instead of compile-time globals:
//   mofreak_main detect <dataset> <out_dir> <video.y8>...                DETECT_MOFREAK
//   mofreak_main bow <dataset> <num_clusters> <num_groups> <svm_path> <file.mofreak>...
// dataset: KTH | TRECVID | UTI2 | HMDB51 | UCF101
#include <stdio.h>
#include <string.h>

#include <chrono>
#include <iostream>
#include <string>

#include "BagOfWordsRepresentation.h"
#include "MoFREAKUtilities.h"

static int datasetId(const std::string& s) {
    const char* names[] = {"KTH", "TRECVID", "HOLLYWOOD", "UTI1", "UTI2", "HMDB51", "UCF101"};
    for (int i = 0; i < 7; i++) if (s == names[i]) return i;
    std::cout << "unknown dataset " << s << std::endl;
    exit(1);
}

int main(int argc, char** argv) {
    if (argc < 4) { std::cout << "usage: see mofreak_main.cpp" << std::endl; return 2; }
    const auto start = std::chrono::steady_clock::now();
    const std::string mode = argv[1];
    const int dset = datasetId(argv[2]);
    if (mode == "detect") {
        MoFREAKUtilities mofreak(dset);
        const std::string out_dir = argv[3];
        for (int i = 4; i < argc; i++) {
            std::string video = argv[i];
            const size_t slash = video.find_last_of('/');
            std::string stem = slash == std::string::npos ? video : video.substr(slash + 1);
            const size_t dot = stem.find_last_of('.');
            if (dot != std::string::npos) stem = stem.substr(0, dot);
            std::cout << "FILE: " << video << std::endl;
            mofreak.computeMoFREAKFromFile(video, out_dir + "/" + stem + ".mofreak", true);
        }
    } else if (mode == "bow" && argc >= 7) {
        BagOfWordsRepresentation bow_rep(atoi(argv[3]), 16, argv[5], atoi(argv[4]), dset);
        bow_rep.intializeBOWMemory(argv[5]);
        for (int i = 6; i < argc; i++) bow_rep.convertFileToBOWFeature(argv[i]);
        bow_rep.writeBOWFeaturesToFiles();
    } else {
        std::cout << "usage: see mofreak_main.cpp" << std::endl;
        return 2;
    }
    const double sec = std::chrono::duration<double>(std::chrono::steady_clock::now() - start).count();
    std::cout << "Took this long: " << sec << " seconds" << std::endl;      // main.cpp:1011
    return 0;
}
