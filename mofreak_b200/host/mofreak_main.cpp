// Command-line driver in the shape of the reference's main.cpp stages (main.cpp:854-924 computeMoFREAKFiles,
// :294-349 computeBOWRepresentation), with paths as arguments instead of compile-time globals:
//   mofreak_main detect <dataset> <out_dir> <video.y8>...                DETECT_MOFREAK
//   mofreak_main bow <dataset> <num_clusters> <num_groups> <svm_path> <file.mofreak>...
//   mofreak_main cluster <dataset> <num_clusters> <num_classes> <svm_path> <seed> <class>:<file.mofreak>...   cluster(), main.cpp:160-213
// dataset: KTH | TRECVID | UTI2 | HMDB51 | UCF101
#include <stdio.h>
#include <string.h>

#include <chrono>
#include <iostream>
#include <string>

#include <vector>

#include "BagOfWordsRepresentation.h"
#include "Clustering.h"
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
    } else if (mode == "cluster" && argc >= 8) {
        const int num_clusters = atoi(argv[3]), num_classes = atoi(argv[4]);
        std::vector<int> classes;
        for (int c = 1; c <= num_classes; c++) classes.push_back(c);
        Clustering clustering(16, num_clusters, 0, num_classes, classes, argv[5]);
        clustering.setAppearanceDescriptor(8, true);
        clustering.setMotionDescriptor(8, true);
        clustering.setSeed((uint32_t)strtoul(argv[6], nullptr, 10));
        for (int c = 1; c <= num_classes; c++) {              // one class at a time, as the reference does to bound memory
            MoFREAKUtilities mofreak(dset);
            std::vector<std::string> files;
            for (int i = 7; i < argc; i++) {
                const std::string a = argv[i];
                const size_t colon = a.find(':');
                if (colon != std::string::npos && atoi(a.substr(0, colon).c_str()) == c) files.push_back(a.substr(colon + 1));
            }
            if (files.empty()) continue;
            for (const std::string& f : files) mofreak.readMoFREAKFeatures(f, 50000 / (int)files.size());
            std::cout << "Building data." << std::endl;
            clustering.buildDataFromMoFREAK(mofreak.getMoFREAKFeatures(), false, false);
            clustering.randomClusters(true);
        }
        clustering.writeClusters(false);
    } else {
        std::cout << "usage: see mofreak_main.cpp" << std::endl;
        return 2;
    }
    const double sec = std::chrono::duration<double>(std::chrono::steady_clock::now() - start).count();
    std::cout << "Took this long: " << sec << " seconds" << std::endl;      // main.cpp:1011
    return 0;
}
