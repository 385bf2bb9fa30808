#include "../include/Logger.h"

static const char *labelOf(typelog t) {
    switch (t) {
        case DEBUG: return "[DEBUG] ";
        case INFO: return "[INFO ] ";
        case WARN: return "[WARN ] ";
        default: return "[ERROR] ";
    }
}
static int colourOf(typelog t) { // ANSI foreground codes: dark gray, light blue, yellow, red
    switch (t) {
        case DEBUG: return 90;
        case INFO: return 94;
        case WARN: return 33;
        default: return 31;
    }
}

Logger::Logger(typelog type) : msglevel(type) {
    if (LOGCFG.headers && enabled()) {
        line << "\033[" << colourOf(type) << "m";
        if (LOGCFG.myRank >= 0) line << "(" << LOGCFG.myRank << ")";
        line << labelOf(type) << "\033[39m";
        opened = true;
    }
}

Logger::~Logger() {
    if (opened) std::cout << line.str() << std::endl;
}
